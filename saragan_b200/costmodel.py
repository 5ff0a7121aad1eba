"""Algorithmic work of the train step: forward conv+linear
FLOPs of G and D per image (2*Cin*Cout*k^3*voxels per layer) and the 4*Gf + 14*Df step model."""
import numpy as np

from .network import num_filters

CONFIGS = {
    # name: (phase, num_phases, base_dim, latent_dim, per-GPU batch)
    "cfg1": dict(phase=3, num_phases=6, base_dim=256, latent_dim=256, batch=4,
                 desc="3D PGAN 'xs' phase 3, 4x16x16 (DxHxW), B=4"),
    "cfg2": dict(phase=5, num_phases=6, base_dim=256, latent_dim=256, batch=2,
                 desc="3D PGAN 'xs' top of progression, 16x64x64, B=2"),
    "cfg3": dict(phase=6, num_phases=7, base_dim=512, latent_dim=512, batch=4,
                 desc="3D PGAN 'small' final phase, 32x128x128, B=4"),
    "cfg4": dict(phase=6, num_phases=8, base_dim=1024, latent_dim=512, batch=4,
                 desc="3D PGAN 'm' final phase, 32x128x128, B=4"),
    "cfg5": dict(phase=7, num_phases=7, base_dim=256, latent_dim=256, batch=2,
                 desc="3D PGAN 'xs' final phase, 64x256x256, B=2"),
}
BASE_SHAPE = (1, 1, 4, 4)


def volume(phase, base_shape=BASE_SHAPE):
    s = 2 ** (phase - 1)
    return (base_shape[1] * s, base_shape[2] * s, base_shape[3] * s)


def conv_layers(kind, phase, num_phases, base_dim):
    """[(name, cin, cout, (D,H,W))] of the 3x3x3 convolutions active at `phase`."""
    f = lambda i: num_filters(i, num_phases, base_dim)
    out = []
    if kind == "g":
        out.append(("g.in", base_dim, base_dim, volume(1)))
        for i in range(2, phase + 1):
            out.append((f"g.b{i}.conv1", f(i), f(i + 1), volume(i)))
            out.append((f"g.b{i}.conv2", f(i + 1), f(i + 1), volume(i)))
    else:
        for i in range(phase, 1, -1):
            out.append((f"d.b{i}.conv1", f(i), f(i), volume(i)))
            out.append((f"d.b{i}.conv2", f(i), f(i - 1), volume(i)))
        out.append(("d.out", base_dim + 1, base_dim, volume(1)))
    return out


def forward_flops_per_image(kind, phase, num_phases, base_dim, latent_dim):
    f = lambda i: num_filters(i, num_phases, base_dim)
    vol0 = int(np.prod(BASE_SHAPE[1:]))
    total = sum(2.0 * ci * co * 27 * int(np.prod(v)) for _, ci, co, v in conv_layers(kind, phase, num_phases, base_dim))
    if kind == "g":
        total += 2.0 * latent_dim * vol0 * base_dim
        total += sum(2.0 * (base_dim if i == 1 else f(i + 1)) * int(np.prod(volume(i))) for i in range(1, phase + 1))
    else:
        total += sum(2.0 * f(i) * int(np.prod(volume(i))) for i in range(1, phase + 1))
        total += 2.0 * vol0 * base_dim * latent_dim + 2.0 * latent_dim
    return total


def step_flops_per_image(phase, num_phases, base_dim, latent_dim, **_):
    gf = forward_flops_per_image("g", phase, num_phases, base_dim, latent_dim)
    df = forward_flops_per_image("d", phase, num_phases, base_dim, latent_dim)
    return 4 * gf + 14 * df


def step_flops_per_image_dict(phase, num_phases, base_dim, latent_dim, **_):
    """4*Gf + 14*Df for the network_dict.py variant: generator block i is f(i-1) -> f(i) -> f(i)
    (network_dict.py:352-356; network.py's is f(i) -> f(i+1)), no minibatch-stddev channel, To/FromRGB at the
    two top levels only."""
    f = lambda i: num_filters(i, num_phases, base_dim)    # noqa: E731
    vol0 = int(np.prod(BASE_SHAPE[1:]))
    vox = lambda i: int(np.prod(volume(i)))               # noqa: E731
    gf = 2.0 * latent_dim * vol0 * base_dim + 2.0 * base_dim * base_dim * 27 * vol0
    gf += sum(2.0 * 27 * (f(i - 1) * f(i) + f(i) * f(i)) * vox(i) for i in range(2, phase + 1))
    gf += 2.0 * f(phase) * vox(phase) + (2.0 * f(phase - 1) * vox(phase - 1) if phase > 1 else 0.0)
    df = sum(2.0 * 27 * (f(i) * f(i) + f(i) * f(i - 1)) * vox(i) for i in range(2, phase + 1))
    df += 2.0 * base_dim * base_dim * 27 * vol0 + 2.0 * vol0 * base_dim * latent_dim + 2.0 * latent_dim
    df += 2.0 * f(phase) * vox(phase) + (2.0 * f(phase - 1) * vox(phase - 1) if phase > 1 else 0.0)
    return 4 * gf + 14 * df
