"""Diffusion mel-refiner train step (BASELINE config 5): `AA_diffusion` under `SpacedDiffusion.training_losses`
(ttts/diffusion/aa_model.py:182-287, ttts/utils/diffusion.py:930-1014, ttts/diffusion/train.py:156-203) on the training tape."""
