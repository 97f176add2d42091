"""Autograd bridge for training (loss.backward() through render_rays).

Round-1 status: the backward kernels (fused MLP backward + compositing backward, SURVEY.md 7
steps 5/7) are not written yet, so requesting gradients fails loudly instead of silently falling
back to a PyTorch path.
"""


def _no_backward(what):
    raise NotImplementedError(
        f"plnerf_b200: {what} was called with gradients enabled, but the sm_100a backward kernels are not "
        "implemented yet (forward/render-only path is complete). Wrap the call in torch.no_grad().")


def mlp_forward_autograd(net, x):
    _no_backward("NeRF.forward")


def render_rays_autograd(*args, **kwargs):
    _no_backward("render_rays")
