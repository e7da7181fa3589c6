"""Restatement of the reference's training losses (test infrastructure).

Follows ``/root/reference/src/train.py``:
  * :95-105  activations: sigmoid / softmax(dim=1) then clamp to [1e-5, 1-1e-5]; bond types are
             soft-maxed over the 6-type axis of ``view(-1, 6, 60, H, W)``; rho = |z|
  * :107-108 atom-centre focal loss, :116-117 bond-centre focal loss (targets in {0, 0.8, 1})
  * :109     atom type (class weights ``train.py:16``), :111 charge, :114 H-count (denominator + 0.1)
  * :119     bond type, :121 rho L1 weighted by the per-(omega,pixel) sum of type targets
  * :124-125 omega focal loss weighted per pixel by the sum of the 60 omega targets
  * :127-135 uncertainty weights ``exp(-s_k) + s_k`` (k = 0,1,2,3,4,7,9) and ``0.5 exp(-s_6) + s_6`` for rho
  * :137     total = plain sum
``multi_gpu_train2.py:152-192`` is identical except that it has no atom-type class weights
(``class_weights=False``).

dtype: the reference allocates the rho / omega targets as float64 (``utils.py:91-92``) so its
losses 4, 5 and the total are float64. ``compute_dtype=None`` keeps whatever
dtypes come in (reference-exact mixing); ``torch.float64`` evaluates everything in double = truth.
"""
from __future__ import annotations

import torch

ATOM_TYPE_WEIGHTS = [1, 0.1, 0.1, 0.1, 1, 1, 1, 1, 1, 10, 10, 10, 10, 10]   # train.py:16
LO, HI = 1e-5, 1 - 1e-5
NAMES = ("atom_targets", "bond_targets", "atom_types", "atom_charges", "bond_types", "bond_rhos",
         "bond_omega_types", "atom_hs")


def _focal_centre(p, t):
    pos = (t == 1).to(p.dtype)
    num = torch.sum(-pos * (1 - p) ** 2 * torch.log(p) - (1 - t) ** 4 * p ** 2 * torch.log(1 - p))
    return num / torch.sum(t == 1)


def losses(outs, targets, s, class_weights: bool = True, compute_dtype=None, n_types: int = 6):
    """outs: 8 logit tensors [B,h,H,W]; targets: (atom_targets, atom_types, atom_charges, atom_hs,
    bond_targets, bond_types[B,6,n_w,H,W], bond_rhos[B,n_w,H,W], bond_omega_types[B,n_w,H,W]);
    s: the 10 uncertainty scalars. Returns (total, dict of the 8 weighted losses, dict of raw losses)."""
    za, zt, zc, zh, zb, zbt, zr, zw = outs
    ta, tt, tc, th, tb, tbt, tr, tw = targets
    if compute_dtype is not None:
        za, zt, zc, zh, zb, zbt, zr, zw = [o.to(compute_dtype) for o in outs]
        ta, tt, tc, th, tb, tbt, tr, tw = [t.to(compute_dtype) for t in targets]
        s = s.to(compute_dtype)
    B, _, H, W = za.shape
    n_w = zw.shape[1]
    pa = torch.clamp(torch.sigmoid(za), LO, HI)
    pt = torch.clamp(torch.softmax(zt, dim=1), LO, HI)
    pc = torch.clamp(torch.softmax(zc, dim=1), LO, HI)
    ph = torch.clamp(torch.softmax(zh, dim=1), LO, HI)
    pb = torch.clamp(torch.sigmoid(zb), LO, HI)
    pbt = torch.clamp(torch.softmax(zbt.view(-1, n_types, n_w, H, W), dim=1), LO, HI)
    pw = torch.clamp(torch.sigmoid(zw), LO, HI)
    rho = torch.abs(zr)

    raw = {}
    raw["atom_targets"] = _focal_centre(pa, ta)
    w = torch.tensor(ATOM_TYPE_WEIGHTS, dtype=pt.dtype).reshape(1, -1, 1, 1) if class_weights else 1.0
    raw["atom_types"] = torch.sum(-w * tt * (1 - pt) ** 2 * torch.log(pt)) / torch.sum(tt)
    raw["atom_charges"] = torch.sum(-tc * (1 - pc) ** 2 * torch.log(pc)) / torch.sum(tc)
    raw["atom_hs"] = torch.sum(-th * (1 - ph) ** 2 * torch.log(ph)) / (torch.sum(th) + 0.1)
    raw["bond_targets"] = _focal_centre(pb, tb)
    raw["bond_types"] = torch.sum(-tbt * (1 - pbt) ** 2 * torch.log(pbt)) / torch.sum(tbt)
    raw["bond_rhos"] = torch.sum(torch.abs(rho - tr) * torch.sum(tbt, dim=1)) / torch.sum(tbt)
    raw["bond_omega_types"] = -torch.sum(
        torch.sum(tw, dim=1, keepdim=True) * ((tw == 1) * ((1 - pw) ** 2) * torch.log(pw)
                                              + (1 - tw) ** 4 * (pw ** 2) * torch.log(1 - pw))) / torch.sum(tw)

    def u(k, half=False):
        return (0.5 if half else 1.0) * torch.exp(-s[k]) + s[k]

    wl = {
        "atom_targets": raw["atom_targets"] * u(0),
        "bond_targets": raw["bond_targets"] * u(1),
        "atom_types": raw["atom_types"] * u(2),
        "atom_charges": raw["atom_charges"] * u(3),
        "bond_types": raw["bond_types"] * u(4),
        "bond_rhos": raw["bond_rhos"] * u(6, half=True),
        "bond_omega_types": raw["bond_omega_types"] * u(7),
        "atom_hs": raw["atom_hs"] * u(9),
    }
    total = (wl["atom_targets"] + wl["bond_targets"] + wl["atom_types"] + wl["atom_charges"]
             + wl["bond_types"] + wl["bond_rhos"] + wl["bond_omega_types"] + wl["atom_hs"])
    return total, wl, raw
