"""Mirror of the reference's `diffusion` package for the denoising hot path (model, optimizer, tools)."""
