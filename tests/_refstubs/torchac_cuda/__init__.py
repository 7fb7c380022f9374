"""Stub of the un-vendored `torchac_cuda` wheel: the three entry points the
reference calls exist but raise (the wheel is not installable here)."""


def _na(*a, **k):
    raise NotImplementedError("torchac_cuda is not available in this container")


encode_fast_new = decode_fast_prefsum = calculate_cdf = _na
