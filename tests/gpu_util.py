"""helpers shared by the GPU parity tests"""
import ctypes as C

import torch

from show_edit_tell_b200 import _lib
from show_edit_tell_b200._lib import EDITNET_FIELDS


def build_module(cls, sd, V, D, A, Fdim, device="cuda"):
    from oracle import synth
    wm = synth.word_map(V)
    mod = cls(wm, decoder_dim=D, caption_features_dim=D, emb_dim=D, attention_dim=A, image_features_dim=Fdim)
    missing, unexpected = mod.load_state_dict(sd, strict=False)
    assert all(k.startswith("caption_encoder.embed.") for k in missing), missing
    assert not unexpected, unexpected
    return mod.to(device), wm


def keep_masks(seed, B, T, Wp, D, R, device="cuda"):
    """the exact dropout keep bits the kernels use, as the oracle's mask dict (CPU tensors)"""
    L = _lib.lib()

    def site(sid, *shape):
        n = 1
        for s in shape:
            n *= s
        out = torch.empty(n, device=device, dtype=torch.float32)
        _lib.check(L.set_dropout_keep_mask(C.c_void_p(out.data_ptr()), n, seed, sid, 0, None))
        torch.cuda.synchronize()
        return out.view(*shape).cpu()

    return {"enc": site(1, B, Wp, D), "emb": site(2, T, B, D), "vis": site(3, T, B, R, D), "fc": site(4, T, B, D)}


def rel_err(a, b):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return float((a - b).abs().max() / max(1e-12, float(b.abs().max())))


def grads_by_key(mod):
    out = {}
    for _, key in EDITNET_FIELDS:
        p = mod.get_parameter(key)
        out[key] = torch.zeros_like(p) if p.grad is None else p.grad.detach().clone()
    return out


def oracle_grads(sd, loss):
    keys = list(sd.keys())
    gs = torch.autograd.grad(loss, [sd[k] for k in keys], allow_unused=True)
    return {k: (g if g is not None else torch.zeros_like(sd[k])) for k, g in zip(keys, gs)}


def compare_grads(mine, ref, tol, label=""):
    """max-normalised comparison; returns list of failing keys with their errors"""
    bad = []
    for k, r in ref.items():
        m = mine[k].detach().cpu()
        # floor: gradients that are analytically zero (the bias in front of a softmax) are pure
        # rounding noise (~1e-10) on both sides
        scale = max(float(r.abs().max()), 1e-5)
        err = float((m - r).abs().max()) / scale
        if not (err < tol):
            bad.append((k, err, scale))
    if bad:
        print("gradient mismatches", label)
        for k, e, s in bad:
            print("   %-55s rel-to-max err %.3e (max |ref| %.3e)" % (k, e, s))
    return bad
