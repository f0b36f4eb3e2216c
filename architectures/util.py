"""Drop-in for the reference's architectures/util.py (freeze helper, util.py:2-10)."""


def freeze_bn_module(m):
    """Put `m` in eval mode if it is a batch-norm layer (class name contains 'BatchNorm'), so it
    normalises with its running statistics; used through `net.apply(freeze_bn_module)`."""
    if 'BatchNorm' in type(m).__name__:
        m.eval()
