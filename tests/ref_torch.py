"""Test helper: the reference's torch op chain restated with plain torch ops (any device).

Follows cachegen_encoder.py:40-61 (torch_quant_vectorized), :76-91 (_split_kv), cachegen_decoder.py:24-35
(do_dequantize) and :182-200 (assembly + cast).  Each op is its own torch kernel, so the fp32 mul and add are
rounded separately exactly as in the reference.  Used for full-size parity on the GPU where the C oracle
would take too long."""
import torch


def quantize(blob_vllm: torch.Tensor, key_bins: torch.Tensor, value_bins: torch.Tensor):
    """blob [L,2,t,H,D] -> (sym int8 [2L,t,C], max_k [L,t,1], max_v [L,t,1])"""
    L, _, t, H, D = blob_vllm.shape
    fp_k, fp_v = torch.unbind(blob_vllm.reshape(L, 2, t, H * D), dim=1)

    def q(bins, x):
        MAX = (bins // 2 - 1)[:, None, None]
        max1 = torch.amax(torch.abs(x), dim=-1, keepdim=True)
        factor = MAX / max1
        xq = torch.round(x * factor + MAX).to(torch.int8)
        return xq, max1

    k, mk = q(key_bins[:L].to(blob_vllm.device), fp_k)
    v, mv = q(value_bins[:L].to(blob_vllm.device), fp_v)
    return torch.cat((k, v), dim=0), mk, mv


def dequantize(sym: torch.Tensor, mk, mv, key_bins, value_bins, H: int, D: int, fmt: str):
    NL, t, C = sym.shape
    L = NL // 2
    key, value = sym.view(torch.uint8).reshape(2, L, t, C).float()

    def dq(x, bins, m):
        Cq = (bins // 2 - 1)[:, None, None]
        x = x - Cq
        x = x / Cq
        return x * m

    key = dq(key, key_bins[:L].to(sym.device), mk)
    value = dq(value, value_bins[:L].to(sym.device), mv)
    blob = torch.stack([key, value]).reshape(2, L, t, H, D)
    if fmt == "vllm":
        return blob.permute(1, 0, 2, 3, 4).to(torch.bfloat16)
    return blob.permute(1, 0, 3, 2, 4).to(torch.float16)


def roundtrip(blob_vllm, key_bins, value_bins, fmt="vllm"):
    L, _, t, H, D = blob_vllm.shape
    sym, mk, mv = quantize(blob_vllm, key_bins, value_bins)
    return dequantize(sym, mk, mv, key_bins, value_bins, H, D, fmt)
