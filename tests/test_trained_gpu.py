"""End-to-end identity on a TRAINED network (north_star: "decoded peak sets ... and the final SMILES must be identical on a
fixed synthetic test set, with any threshold-boundary ties listed explicitly").

The reference ships no weights, and random-init logits are nearly flat, so the test first trains the network on labelled
pseudo-molecule drawings (oracle/synth.pseudo_molecules + rasterise_targets, utils.py:83-228) with the PRODUCT training
step (abcnet_b200.TrainStep: CUDA forward / losses / backward + Adam, one CUDA graph). Then, on the fixed set of N_IMG = 256
images, the product inference path (bf16 activations, fp32 logits) + CUDA decode is compared with the oracle: fp32 CPU forward
of src/unet.py's graph (oracle/unet_ref) + the restated decode statements of img2smiles.py:62-193.

What is asserted, in this order
  1. DECODE IS EXACT ON THE PRODUCT'S OWN LOGITS: for every image the CUDA decoder's records equal, bit for bit (positions,
     classes, omega survivors, order, rho bits), the oracle decode applied to the product's fp32 logits. Hence every difference
     to the reference below is a consequence of logit error alone, never of the decode kernel.
  2. logits (bf16 activations through 45 layers vs fp32; trained weights with logit ranges of 20..170): per output map
     max-abs error <= 0.08 * max|ref| + 0.05 and relative L2 error <= 5e-2.
  3. every record that differs between decode(reference logits) and decode(product logits) is listed with the comparison that
     flipped (threshold, 3x3 / 3-bin NMS neighbour, antipodal omega bin, arg-max runner-up), its margin in the fp32 reference
     logits and the LOCAL error = |ours - ref| summed over exactly the two logits of that comparison. A flip is a boundary tie iff
     margin <= local error; the old whole-map bound (2 x max|err| of the map) is gone.
  4. for EVERY image -- differing ones included -- both record sets go through the reference's host assembly (assemble_ref):
     MOL-block text equality and molecular-graph equality (assemble_ref.molecule_graph: an omega / omega+30 flip of an undirected
     bond only swaps that bond's two end atoms; a one-pixel move of an atom peak on a two-pixel plateau changes a drawing
     coordinate but not the topology) are counted and reported; a graph change must come from a listed boundary case.
  5. error budget per stage: the product's own trunk fed to fp32 heads separates trunk error from head error (hidden-map
     rounding, bf16 conv1 / conv2 weights), reported per decision head.
The report is written to gpurun_out/trained_parity_report.json (copied to profiles/ when the run is recorded).
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import assemble_ref, decode_ref, synth, unet_ref

pytestmark = pytest.mark.gpu
HEADS = list(unet_ref.V2_HEADS)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_IMG = int(os.environ.get("ABCNET_TRAINED_IMAGES", "256"))
BATCH, CHUNK = 16, 32
STEPS, LR, LR_LATE = int(os.environ.get("ABCNET_TRAINED_STEPS", "4000")), 6e-4, 2.5e-4   # late = train.py:55
THR = -1.0


def _targets_on_gpu(labels):
    """Dense targets of all images, rasterised chunk-wise on the host (oracle rules) and kept in HBM (41.7 MB per image)."""
    parts = None
    for c0 in range(0, len(labels), CHUNK):
        tg = synth.rasterise_targets(labels[c0:c0 + CHUNK], 128, 128)
        if parts is None:
            parts = [torch.empty((len(labels),) + t.shape[1:], dtype=torch.from_numpy(t).dtype, device="cuda") for t in tg]
        for dst, t in zip(parts, tg):
            dst[c0:c0 + CHUNK].copy_(torch.from_numpy(t))
    return parts


def _train(imgs, targets):
    import abcnet_b200
    torch.manual_seed(0)
    m = abcnet_b200.UNet(1, HEADS).cuda()
    m.load_state_dict(unet_ref.make_state_dict(seed=11, variant="W0"))
    m.train()
    lr = torch.tensor(LR, device="cuda")                   # a device tensor: can be changed between CUDA-graph replays
    opt = abcnet_b200.make_optimizer(m, lr=lr, capturable=True)
    step = abcnet_b200.TrainStep(m, opt, class_weights=True, use_graph=True)
    x = torch.from_numpy(imgs).cuda()
    nb = len(imgs) // BATCH
    curve = []
    for it in range(STEPS):
        b = it % nb
        if it == (STEPS * 5) // 8:
            lr.fill_(LR_LATE)
        sl = slice(b * BATCH, (b + 1) * BATCH)
        loss = step(x[sl].contiguous(), [t[sl].contiguous() for t in targets])
        if it % 200 == 0 or it == STEPS - 1:
            curve.append((it, float(loss.item())))
    torch.cuda.synchronize()
    m.eval()
    return m, curve


# ---------------------------------------------------------------------------- which comparison flipped?
def _flips(cmps):
    """cmps: (name, ref_a, ref_b, our_a, our_b, strict) for predicates 'a > b' (strict) or 'a >= b'; b may be a constant
    (then our_b == ref_b). Returns the flipped predicates as dicts with margin and local error."""
    out = []
    for name, ra, rb, oa, ob, strict in cmps:
        pr = (ra > rb) if strict else (ra >= rb)
        po = (oa > ob) if strict else (oa >= ob)
        if pr != po:
            out.append(dict(cmp=name, margin=float(abs(float(ra) - float(rb))),
                            local_err=float(abs(float(oa) - float(ra)) + abs(float(ob) - float(rb)))))
    return out


def _centre_cmps(zr, zo, x, y):
    H, W = zr.shape
    c = [("thr", zr[x, y], THR, zo[x, y], THR, True)]
    for i in range(max(x - 1, 0), min(x + 2, H)):
        for j in range(max(y - 1, 0), min(y + 2, W)):
            if (i, j) != (x, y):
                c.append((f"nms({i - x},{j - y})", zr[x, y], zr[i, j], zo[x, y], zo[i, j], False))
    return c


def _omega_cmps(cr, co, w):
    """The comparisons that decide whether bin w of an omega column is emitted (img2smiles.py:74-80, :143-158)."""
    n = len(cr)
    h = n // 2
    c = [("thr", cr[w], THR, co[w], THR, True)]
    for k in ((w - 1) % n, (w + 1) % n):
        c.append((f"nms(bin {k})", cr[w], cr[k], co[w], co[k], False))
    if w <= h - 2:
        others, strict = (w + h - 1, w + h), False          # dropped if z_w <  max(...): survives on z_w >= each
    elif w == h - 1:
        others, strict = (n - 2, 0), False
    elif w == h:
        others, strict = (0, n - 1), True                   # dropped if z_w <= ...: survives on z_w > each
    else:
        others, strict = (w - h - 1, w - h), True
    for k in others:
        c.append((f"antipode(bin {k})", cr[w], cr[k], co[w], co[k], strict))
    return c


def _argmax_cmps(vr, vo, name):
    a, b = int(np.argmax(vr)), int(np.argmax(vo))
    # the reference prefers a; ours prefers b: 'v[a] beats v[b]' (first maximum wins ties: strict iff b < a)
    return [(f"{name} {a} vs {b}", vr[a], vr[b], vo[a], vo[b], b < a)]


def _explain(j, R, O, ra, rb, oa, ob):
    """All record differences of image j between decode(R) and decode(O), each with its flipped comparison(s)."""
    d = []
    A_r, A_o = {(a[0], a[1]): a for a in ra.tolist()}, {(a[0], a[1]): a for a in oa.tolist()}
    for pos in sorted(set(A_r) ^ set(A_o)):
        d.append(dict(image=j, kind="atom peak", pos=pos, in_ref=pos in A_r, flips=_flips(_centre_cmps(R[0][0], O[0][0], *pos))))
    for pos in sorted(set(A_r) & set(A_o)):
        for name, col, head in (("atom type", 2, 1), ("atom charge", 3, 2), ("atom hs", 4, 3)):
            if A_r[pos][col] != A_o[pos][col]:
                d.append(dict(image=j, kind=name, pos=pos, ref=A_r[pos][col], ours=A_o[pos][col],
                              flips=_flips(_argmax_cmps(R[head][:, pos[0], pos[1]], O[head][:, pos[0], pos[1]], name))))
    B_r = {(b[0], b[1], b[2]): b[3] for b in rb.tolist()}
    B_o = {(b[0], b[1], b[2]): b[3] for b in ob.tolist()}
    P_r = decode_ref._peaks2d(R[4][0], THR)
    P_o = decode_ref._peaks2d(O[4][0], THR)
    for key in sorted(set(B_r) ^ set(B_o)):
        x_, y_, w = key
        if bool(P_r[x_, y_]) != bool(P_o[x_, y_]):
            d.append(dict(image=j, kind="bond peak", pos=(x_, y_), omega=w, in_ref=key in B_r,
                          flips=_flips(_centre_cmps(R[4][0], O[4][0], x_, y_))))
        else:
            d.append(dict(image=j, kind="bond omega", pos=(x_, y_), omega=w, in_ref=key in B_r,
                          flips=_flips(_omega_cmps(R[7][:, x_, y_], O[7][:, x_, y_], w))))
    n_w = R[7].shape[0]
    for key in sorted(set(B_r) & set(B_o)):
        if B_r[key] != B_o[key]:
            x_, y_, w = key
            vr = R[5].reshape(-1, n_w, *R[5].shape[1:])[:, w, x_, y_]
            vo = O[5].reshape(-1, n_w, *O[5].shape[1:])[:, w, x_, y_]
            d.append(dict(image=j, kind="bond type", pos=(x_, y_), omega=w, ref=B_r[key], ours=B_o[key],
                          flips=_flips(_argmax_cmps(vr, vo, "bond type"))))
    return d


def test_trained_network_end_to_end_identity():
    import abcnet_b200
    imgs, labels = synth.pseudo_molecules(7, N_IMG, 512, 512)
    targets = _targets_on_gpu(labels)
    model, curve = _train(imgs, targets)
    del targets
    torch.cuda.empty_cache()
    print("training loss curve:", curve)
    assert curve[-1][1] < 0.2 * curve[0][1], "training did not converge; the decode comparison would be meaningless"
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    sd_dev = {k: v.detach().clone() for k, v in model.state_dict().items()}

    torch.set_num_threads(os.cpu_count() or 1)
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False       # fp32 hybrid heads below: true fp32
    dec = abcnet_b200.PeakDecoder(CHUNK, atom_cap=2048, bond_cap=8192)
    pipe = abcnet_b200.SparseHeadsPipeline(model, CHUNK, peak_cap=256, bond_cap=8192)
    # the same trained weights in the decision-stable inference mode (fp16 activations / weights: 8 x smaller rounding error)
    model16 = abcnet_b200.UNet(1, HEADS, act_dtype="fp16").cuda().eval()
    model16.load_state_dict(model.state_dict())
    dec16 = abcnet_b200.PeakDecoder(CHUNK, atom_cap=2048, bond_cap=8192)
    f16 = dict(identical_record_images=0, identical_molblock_images=0, identical_topology_images=0, differing_decisions=0,
               logit_max_abs_err=[0.0] * 8)
    report = dict(images=N_IMG, train_steps=STEPS, loss_curve=curve, differences=[], identical_record_images=0,
                  identical_molblock_images=0, identical_graph_images=0, identical_topology_images=0, molecules_compared=0, labelled_atoms=0, found_atoms=0,
                  ref_atom_peaks=0, ref_bond_records=0, images_with_differences=[], graph_changes=[])
    err = np.zeros(8)
    scale = np.zeros(8)
    num = np.zeros(8)
    den = np.zeros(8)
    budget = {k: dict(trunk=0.0, heads=0.0, total=0.0) for k in (0, 4, 7)}
    unexplained = []
    for c0 in range(0, N_IMG, CHUNK):
        xs = torch.from_numpy(imgs[c0:c0 + CHUNK]).cuda()
        n = xs.shape[0]
        # product: dense maps in the fused layout -> CUDA decode; reference-format logits for the comparison
        recs = dec(model.infer(xs, layout="p8f"), thr=THR)
        blocks = dec.molblocks(n)
        ours_t = model(xs)
        ours = [o.float().cpu().numpy() for o in ours_t]
        # the opt-in sparse-heads path on the same trained network and images: identical records and MOL blocks
        sparse = pipe.fetch(pipe.launch(xs, thr=THR)) if n == CHUNK else None
        if sparse is not None:
            for j, ((da, db, dn), (sa, sb, sn)) in enumerate(zip(recs, sparse)):
                assert dn == sn and np.array_equal(da, sa) and np.array_equal(db, sb), f"image {c0 + j}: sparse-heads records differ"
            assert pipe.molblocks(n) == blocks
        # oracle: fp32 CPU forward
        with torch.no_grad():
            ref = [o.numpy() for o in unet_ref.forward(torch.from_numpy(imgs[c0:c0 + CHUNK]), sd)]
        for i in range(8):
            err[i] = max(err[i], float(np.abs(ours[i] - ref[i]).max()))
            scale[i] = max(scale[i], float(np.abs(ref[i]).max()))
            num[i] += float(((ours[i] - ref[i]).astype(np.float64) ** 2).sum())
            den[i] += float((ref[i].astype(np.float64) ** 2).sum())
        # error budget: the product's trunk (bf16) through fp32 heads (cuDNN fp32, TF32 off) isolates the trunk's share
        with torch.no_grad():
            hyb = unet_ref.heads_forward(model.activation("k2").float(), sd_dev)
        for k in budget:
            h = hyb[k].cpu().numpy()
            budget[k]["trunk"] = max(budget[k]["trunk"], float(np.abs(h - ref[k]).max()))
            budget[k]["heads"] = max(budget[k]["heads"], float(np.abs(ours[k] - h).max()))
            budget[k]["total"] = max(budget[k]["total"], float(np.abs(ours[k] - ref[k]).max()))
        recs16 = dec16(model16.infer(xs, layout="p8f"), thr=THR)
        ours16 = [o.float().cpu().numpy() for o in model16(xs)]
        for i in range(8):
            f16["logit_max_abs_err"][i] = max(f16["logit_max_abs_err"][i], float(np.abs(ours16[i] - ref[i]).max()))
        for jj in range(n):
            j = c0 + jj
            R = [r[jj] for r in ref]
            O = [o[jj] for o in ours]
            ra, (rb, rrho) = decode_ref.decode_records(R, THR, "nms")
            oa, (ob, orho) = decode_ref.decode_records(O, THR, "nms")
            atoms, bonds, nbp = recs[jj]
            ga = np.stack([atoms["x"], atoms["y"], atoms["type"], atoms["charge"], atoms["hs"]], -1).astype(np.int32).reshape(-1, 5)
            gb = np.stack([bonds["x"], bonds["y"], bonds["omega"], bonds["type"]], -1).astype(np.int32).reshape(-1, 4)
            # 1. the CUDA decoder is exact on the product's own logits
            assert np.array_equal(ga, oa) and np.array_equal(gb, ob), f"image {j}: CUDA decode != oracle decode on the SAME logits"
            assert np.array_equal(bonds["rho"].view(np.uint32), orho.view(np.uint32)), f"image {j}: rho bits"
            report["ref_atom_peaks"] += len(ra)
            report["ref_bond_records"] += len(rb)
            lab = {(a[0] // 4, a[1] // 4) for a in labels[j]["atoms"]}
            report["labelled_atoms"] += len(lab)
            report["found_atoms"] += sum(1 for (px, py) in lab if any(abs(px - q[0]) <= 1 and abs(py - q[1]) <= 1 for q in ra))
            # 4. molecule level, for every image
            L_ref = decode_ref.records_to_lists(ra, (rb, rrho)) if (len(ra) and len(rb)) else None
            L_our = abcnet_b200.records_to_lists(atoms, bonds, nbp)
            mol_r = assemble_ref.records_to_molblock(L_ref) if L_ref is not None else None
            mol_o = assemble_ref.records_to_molblock(L_our) if L_our is not None else None
            g_r = assemble_ref.molecule_graph(L_ref) if L_ref is not None else None
            g_o = assemble_ref.molecule_graph(L_our) if L_our is not None else None
            t_r = assemble_ref.molecule_graph(L_ref, with_positions=False) if L_ref is not None else None
            t_o = assemble_ref.molecule_graph(L_our, with_positions=False) if L_our is not None else None
            report["identical_topology_images"] += int(t_r == t_o)
            assert mol_o == blocks[jj], f"image {j}: native assembler text != reference assembly of the same records"
            report["molecules_compared"] += int(mol_r is not None)
            # fp16 mode on the same image: records / MOL text / topology against the same reference
            a16, b16, n16 = recs16[jj]
            ga16 = np.stack([a16["x"], a16["y"], a16["type"], a16["charge"], a16["hs"]], -1).astype(np.int32).reshape(-1, 5)
            gb16 = np.stack([b16["x"], b16["y"], b16["omega"], b16["type"]], -1).astype(np.int32).reshape(-1, 4)
            same16 = np.array_equal(ga16, ra) and np.array_equal(gb16, rb)
            f16["identical_record_images"] += int(same16)
            L16 = abcnet_b200.records_to_lists(a16, b16, n16)
            f16["identical_molblock_images"] += int((assemble_ref.records_to_molblock(L16) if L16 is not None else None) == mol_r)
            f16["identical_topology_images"] += int((assemble_ref.molecule_graph(L16, with_positions=False) if L16 is not None else None) ==
                                                    (assemble_ref.molecule_graph(L_ref, with_positions=False) if L_ref is not None else None))
            if not same16:
                O16 = [o[jj] for o in ours16]
                oa16, (ob16, _) = decode_ref.decode_records(O16, THR, "nms")
                assert np.array_equal(ga16, oa16) and np.array_equal(gb16, ob16), f"image {j}: fp16 mode: CUDA decode != oracle decode on the same logits"
                d16 = _explain(j, R, O16, ra, rb, oa16, ob16)
                f16["differing_decisions"] += len(d16)
                for item in d16:
                    if not (item["flips"] and all(f["margin"] <= f["local_err"] for f in item["flips"])):
                        unexplained.append(dict(item, mode="fp16"))
            same_rec = np.array_equal(ga, ra) and np.array_equal(gb, rb)
            report["identical_record_images"] += int(same_rec)
            report["identical_molblock_images"] += int(mol_r == mol_o)
            report["identical_graph_images"] += int(g_r == g_o)
            if same_rec:
                assert mol_r == mol_o, f"image {j}: identical records but different MOL block (rho dependence)"
                if len(rb):
                    assert np.abs(bonds["rho"] - rrho).max() <= 0.08 * scale[6] + 0.05
                continue
            # 3. list and explain every difference
            d = _explain(j, R, O, ra, rb, oa, ob)
            assert d, f"image {j}: records differ only in order"
            for item in d:
                item["explained"] = bool(item["flips"]) and all(f["margin"] <= f["local_err"] for f in item["flips"])
                if not item["explained"]:
                    unexplained.append(item)
            report["differences"] += d
            report["images_with_differences"].append(dict(image=j, n=len(d), molblock_identical=bool(mol_r == mol_o),
                                                          graph_identical=bool(g_r == g_o), topology_identical=bool(t_r == t_o),
                                                          kinds=sorted({i["kind"] for i in d})))
            if g_r != g_o:
                report["graph_changes"].append(dict(image=j, kinds=sorted({i["kind"] for i in d})))
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    rel = np.sqrt(num / np.maximum(den, 1e-30))
    report.update(logit_max_abs_err=err.tolist(), logit_scale=scale.tolist(), logit_rel_l2=rel.tolist(),
                  error_budget_max_abs={str(k): v for k, v in budget.items()}, fp16_mode=f16)
    margins = [f["margin"] for it in report["differences"] for f in it["flips"]]
    report["flip_margin_max"] = max(margins) if margins else 0.0
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "trained_parity_report.json"), "w") as f:
        json.dump(report, f, indent=1, default=lambda o: o.item() if hasattr(o, "item") else str(o))
    print("logit max-abs error per map:", [round(e, 4) for e in err], "scale:", [round(s, 2) for s in scale],
          "rel L2:", [round(r, 4) for r in rel])
    print("error budget (max-abs, decision heads): ", budget)
    print(f"images {N_IMG}: identical records {report['identical_record_images']}, identical MOL text {report['identical_molblock_images']}, "
          f"identical molecular graph {report['identical_graph_images']} (topology without drawing coordinates: "
          f"{report['identical_topology_images']}); molecules {report['molecules_compared']}; reference atom peaks "
          f"{report['ref_atom_peaks']} (labelled {report['labelled_atoms']}, found {report['found_atoms']}), bond records "
          f"{report['ref_bond_records']}; listed boundary cases: {len(report['differences'])}, largest flipped margin {report['flip_margin_max']:.4f}")
    print(f"fp16 activation mode on the same weights: identical records {f16['identical_record_images']}, identical MOL text "
          f"{f16['identical_molblock_images']}, identical topology {f16['identical_topology_images']}, differing decisions "
          f"{f16['differing_decisions']} (bf16: {len(report['differences'])}); logit max-abs error "
          f"{[round(e, 4) for e in f16['logit_max_abs_err']]}")
    for item in report["differences"][:40]:
        print("  boundary case:", item)
    for i in range(8):
        assert err[i] <= 0.08 * scale[i] + 0.05, f"map {i}: max abs err {err[i]} (scale {scale[i]})"
        assert rel[i] <= 0.05, f"map {i}: rel L2 {rel[i]}"
    assert not unexplained, f"differences that are NOT threshold / tie boundary cases: {unexplained[:5]}"
    assert report["ref_atom_peaks"] >= 4 * N_IMG and report["found_atoms"] >= 0.7 * report["labelled_atoms"], "degenerate network"
    assert report["identical_record_images"] >= (2 * N_IMG) // 3
    assert report["identical_graph_images"] >= (9 * N_IMG) // 10
    # the decision-stable mode must be at least as close to the reference as bf16, in logits and in decisions
    assert all(a <= b + 1e-6 for a, b in zip(f16["logit_max_abs_err"], err.tolist())), (f16["logit_max_abs_err"], err.tolist())
    assert f16["identical_record_images"] >= report["identical_record_images"]
    assert f16["differing_decisions"] <= len(report["differences"])


def test_eval_cache_follows_graph_training():
    """a CUDA-graph replay updates weights and BatchNorm running statistics behind torch's version
    counters; eval -> train (replays) -> eval must run on the NEW weights (train.py:89,218 alternates per epoch)."""
    import abcnet_b200
    imgs, labels = synth.pseudo_molecules(3, 8, 256, 256)
    tg = [torch.from_numpy(t).cuda() for t in synth.rasterise_targets(labels, 64, 64)]
    x = torch.from_numpy(imgs).cuda()
    torch.manual_seed(0)
    m = abcnet_b200.UNet(1, HEADS).cuda()
    m.load_state_dict(unet_ref.make_state_dict(seed=5, variant="W1"))
    opt = abcnet_b200.make_optimizer(m, lr=1e-3)
    step = abcnet_b200.TrainStep(m, opt, use_graph=True)
    pipe = abcnet_b200.SparseHeadsPipeline(m, 8, peak_cap=1024, bond_cap=8192)
    m.train()
    for _ in range(2):
        step(x, tg)
    m.eval()
    first = [o.clone() for o in m(x)]
    pipe.launch(x)                                          # caches a pack derived from the current weights
    m.train()
    for _ in range(5):
        step(x, tg)                                         # pure graph replays: no _version bump anywhere
    m.eval()
    second = m(x)
    fresh = abcnet_b200.UNet(1, HEADS).cuda().eval()
    fresh.load_state_dict(m.state_dict())
    want = fresh(x)
    assert any((a - b).abs().max().item() > 1e-3 for a, b in zip(first, second)), "training did not change the outputs"
    for i, (a, b) in enumerate(zip(second, want)):
        assert torch.equal(a, b), f"head {i}: eval after graph training used stale packed weights"
    # the sparse pipeline re-derives its centre pack as well
    dec = abcnet_b200.PeakDecoder(8, atom_cap=4096, bond_cap=16384)
    dense = dec(fresh.infer(x, layout="p8f"))
    try:
        sp = pipe.fetch(pipe.launch(x))
    except RuntimeError:
        return                                              # more peaks than peak_cap on this barely trained net: nothing to compare
    for (da, db, dn), (sa, sb, sn) in zip(dense, sp):
        assert dn == sn and np.array_equal(da, sa) and np.array_equal(db, sb)
