"""Drop-in for the reference's ``pgan_pytorch/loss.py`` (same names and arguments)."""
import torch

from . import ops


def wasserstein_loss(y_pred):
    """loss.py:3-4."""
    return y_pred.mean()


def compute_gradient_penalty(discriminator, real_samples, fake_samples, alpha,
                             gradient_penalty_weight=10, random_uniform=None):
    """loss.py:7-27: WGAN-GP, target norm 1.  ``random_uniform`` (B,1,1,1,1) may be supplied
    to replay a fixed draw (parity tests); by default it is drawn like the reference does."""
    if random_uniform is None:
        random_uniform = torch.rand(real_samples.shape[0], 1, 1, 1, 1)
    random_uniform = random_uniform.to(real_samples.device)
    interpolates = ops.interpolate(real_samples, fake_samples, random_uniform).requires_grad_(True)
    d_interpolates = discriminator(interpolates, alpha)
    ones = torch.ones_like(d_interpolates)
    # only the input gradient is wanted here (only_inputs=True): skip every layer's wgrad
    with ops.no_weight_gradients():
        gradients = torch.autograd.grad(
            outputs=d_interpolates,
            inputs=interpolates,
            grad_outputs=ones,
            create_graph=True,
            retain_graph=True,
            only_inputs=True,
        )[0]
    norms = ops.RowNorm.apply(gradients)
    gradient_penalty = ((norms - 1) ** 2).mean()
    return gradient_penalty * gradient_penalty_weight
