"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Test infrastructure (see oracle/simt_oracle.py header).  Runs only in the
build container, where the reference is mounted read-only at /root/reference;
the GPU box never sees that path, so the vectors are committed.

What is imported from the reference, untouched:
  * utils/loss.py            -> CrossEntropy2d
  * model/deeplab_multi.py   -> sig_NTM, sig_W  (their forward() hard-calls
    .cuda(); on this CPU-only box ``torch.Tensor.cuda`` is shimmed to identity,
    and cwd is the reference's tools/ because __init__ opens
    '../ClassDist/ClassDist_bapa.npy' by relative path)
  * tools/compute_iou.py     -> fast_hist, per_class_iu, label_mapping
The head composition (tools/trainV2_simt.py:371-372,402-409) and the two
one-line histograms of compute_ConfusionMatrix.py:54-56 /
compute_ClassDistribution.py:52-54 cannot be imported (those modules run
argparse / import ttach, matplotlib at import time), so for them the script
executes the cited source lines' op sequence through the imported
CrossEntropy2d / torch / numpy directly.

Usage:  python oracle/make_golden.py   (rewrites tests/golden/)
"""
import json
import os
import sys
import zlib

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, os.path.dirname(HERE))


def _import_reference():
    os.chdir(os.path.join(REF, "tools"))
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "tools"))
    torch.Tensor.cuda = lambda self, *a, **k: self      # CPU-only container
    torch.nn.Module.cuda = lambda self, *a, **k: self
    from utils.loss import CrossEntropy2d
    from model.deeplab_multi import sig_NTM, sig_W
    import compute_iou
    return CrossEntropy2d, sig_NTM, sig_W, compute_iou


def ref_head(CrossEntropy2d, logits, T, labels, size, dtype):
    """tools/trainV2_simt.py:301,371-372,402-409 executed with the reference's own modules."""
    import torch.nn as nn
    B, CK = logits.shape[:2]
    H, W = size
    C = T.shape[1]
    interp_target = nn.Upsample(size=(H, W), mode="bilinear", align_corners=True)   # :301
    Tseg_loss = CrossEntropy2d(is_softmax=False)                                    # :304
    lg = logits.detach().to(dtype).clone().requires_grad_(True)
    Tt = T.detach().to(dtype).clone().requires_grad_(True)
    pred = interp_target(lg)                                                         # :371
    pred = torch.softmax(interp_target(pred), dim=1).permute(0, 2, 3, 1).contiguous().view(-1, CK)  # :402
    pred = torch.mm(pred, Tt).view(B, H, W, C).permute(0, 3, 1, 2)                   # :403
    loss = Tseg_loss(pred, labels.long())                                            # :408
    loss.backward()                                                                  # :428
    return loss.detach(), lg.grad.detach(), Tt.grad.detach()


HEAD_CASES = [
    # name,            B, K,  h,  w,   H,   W, coherent, ignore_frac
    ("head_small_u",   2, 0,  5,  9,  32,  64, False, 0.10),
    ("head_small_r",   2, 0,  5,  9,  32,  64, True,  0.10),
    ("head_openset4",  1, 4,  9, 17,  64, 128, True,  0.10),
    ("head_openset15", 1, 15, 5,  9,  33,  65, True,  0.25),
    ("head_odd",       3, 0,  7,  6,  41,  29, False, 0.00),
    ("head_identity",  1, 0, 12, 20,  12,  20, False, 0.10),
    ("head_down",      1, 0, 17, 33,   8,  16, False, 0.10),
    ("head_row1",      1, 0,  1,  9,   1,  64, True,  0.10),
    ("head_cfg1_tile", 1, 0, 17, 33, 128, 256, True,  0.10),
]


def close(x, y, what, tol):
    x, y = torch.as_tensor(x).double(), torch.as_tensor(y).double()
    err, ref = float((x - y).norm()), float(y.norm())
    assert err <= tol * max(ref, 1e-30), f"{what}: restatement differs from the reference ({err:.3e} vs norm {ref:.3e})"


def _reference_placeholder_loss():
    """The reference's own ``Placeholder_loss`` (tools/trainV2_simt.py:202-230), compiled from its source file.
    The script cannot be imported (it parses the command line and pulls in datasets / matplotlib at import
    time), so the one FunctionDef is taken out of the file with ``ast`` and executed unmodified."""
    import ast
    import types
    path = os.path.join(REF, "tools", "trainV2_simt.py")
    tree = ast.parse(open(path).read(), path)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "Placeholder_loss")
    ns = {"torch": torch, "args": types.SimpleNamespace(num_classes=19, open_classes=0, lambda_Place=0.1)}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns["Placeholder_loss"], ns["args"]


def gen_placeholder(O, manifest):
    import torch.nn as nn
    ref_fn, args = _reference_placeholder_loss()
    cases = [  # name, B, K, h, w, H, W, thres, logit scale
        ("place_K4", 2, 4, 9, 17, 64, 128, 0.8, 3.0),
        ("place_K15", 1, 15, 5, 9, 33, 65, 0.8, 4.0),
        ("place_K4_nothres", 1, 4, 7, 6, 41, 29, None, 2.0),
    ]
    for name, B, K, h, w, H, W, thres, sc in cases:
        CK = 19 + K
        args.num_classes, args.open_classes, args.lambda_Place = 19, K, 0.1           # :63 default
        g = torch.Generator().manual_seed(zlib.crc32(name.encode()) % 10007)
        lo = sc * torch.randn(B, CK, h, w, generator=g)
        # make a good share of pixels confidently known-class, some confidently open-set, some below the threshold
        boost = torch.randint(0, CK, (B, 1, h, w), generator=g)
        lo = lo + 6.0 * torch.nn.functional.one_hot(boost[:, 0], CK).permute(0, 3, 1, 2) * \
            (torch.rand(B, 1, h, w, generator=g) < 0.7)
        interp = nn.Upsample(size=(H, W), mode="bilinear", align_corners=True)           # :301
        out = {"logits": lo.numpy(), "size": np.array([H, W]), "thres": np.float64(-1.0 if thres is None else thres),
               "lambda_place": np.float64(0.1), "K": np.int64(K)}
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            lg = lo.to(dt).clone().requires_grad_(True)
            loss = ref_fn(interp(lg), 19, K, thres=thres)                                # :371,398
            loss.backward()
            l_o, g_o = O.placeholder_fwd_bwd(lo, (H, W), 19, K, thres, 0.1, dt)
            close(l_o, loss.detach(), name + " loss", tol=1e-6)
            close(g_o, lg.grad, name + " grad", tol=1e-6)
            out[f"loss_{tag}"] = loss.detach().numpy()
            out[f"dlogits_{tag}"] = lg.grad.numpy()
        up = interp(lo)
        am = up.argmax(1)
        out["n_valid"] = np.int64(((am < 19) & ((torch.softmax(up, 1).max(1)[0] > thres) if thres is not None
                                                else torch.ones_like(am, dtype=torch.bool))).sum())
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        manifest["cases"].append(name)


def gen_wfit(sig_NTM, sig_W, O, manifest):
    # ---- inner W optimisation (a8), trainV2_simt.py:271-280,317-339 with the reference's modules + torch Adam ----
    for K in (4, 15):
        CK = 19 + K
        torch.manual_seed(777 + K)
        ntm = [sig_NTM(19, K), sig_NTM(19, K)]
        wm = [sig_W(19, K), sig_W(19, K)]
        lr = 2.5e-4                                                                      # LEARNING_RATE_T (:48)
        opt_t = [torch.optim.Adam(m.parameters(), lr=lr, weight_decay=0) for m in ntm]  # :271,274
        opt_w = [torch.optim.Adam(m.parameters(), lr=lr, weight_decay=0) for m in wm]   # :277,280
        loss_mse = torch.nn.MSELoss(reduction="sum")
        zeros = torch.zeros(CK, 19)
        # oracle twin, from the same initial state
        o_ntm = [m.NTM.detach().clone() for m in ntm]
        o_w = [m.weight.detach().clone() for m in wm]
        o_st = [dict(m=torch.zeros(CK, CK), v=torch.zeros(CK, CK), step=0) for _ in wm]
        cd = np.load(os.path.join(REF, "ClassDist", "ClassDist_bapa.npy"))
        out = dict(ntm1=o_ntm[0].numpy().copy(), ntm2=o_ntm[1].numpy().copy(), w1_init=o_w[0].numpy().copy(),
                   w2_init=o_w[1].numpy().copy(), lr=np.float64(lr), rounds=np.int64(10))
        for outer in range(2):                      # two outer iterations: the second starts from Adam step 10
            for o in opt_t + opt_w:
                o.zero_grad()                                                            # :317-320
            ref_losses = []
            for it in range(10):                                                         # :327
                T1, T2, W1, W2 = ntm[0](), ntm[1](), wm[0](), wm[1]()                    # :329-332
                opt_w[0].zero_grad(); opt_w[1].zero_grad()                               # :333-334
                NTM_loss = loss_mse(W1.mm(T1), zeros) + loss_mse(W2.mm(T2), zeros)       # :336
                NTM_loss.backward(retain_graph=True)                                     # :337
                opt_w[0].step(); opt_w[1].step()                                         # :338-339
                ref_losses.append(NTM_loss.detach())
            g_o, l_o = O.w_fit_loop(o_ntm, o_w, o_st, lr, cd, 19, K, rounds=10)
            for i in range(2):
                stt = opt_w[i].state[wm[i].weight]
                assert int(stt["step"]) == o_st[i]["step"] == 10 * (outer + 1)
                close(o_w[i], wm[i].weight.detach(), f"wfit K{K} weight{i} outer{outer}", tol=1e-6)
                close(o_st[i]["m"], stt["exp_avg"], f"wfit K{K} exp_avg{i}", tol=1e-5)
                close(o_st[i]["v"], stt["exp_avg_sq"], f"wfit K{K} exp_avg_sq{i}", tol=1e-5)
                close(g_o[i], ntm[i].NTM.grad, f"wfit K{K} ntm grad{i}", tol=1e-5)
                out[f"w{i + 1}_after{outer}"] = wm[i].weight.detach().numpy().copy()
                out[f"m{i + 1}_after{outer}"] = stt["exp_avg"].numpy().copy()
                out[f"v{i + 1}_after{outer}"] = stt["exp_avg_sq"].numpy().copy()
                out[f"ntm_grad{i + 1}_outer{outer}"] = ntm[i].NTM.grad.numpy().copy()
            close(l_o, torch.stack(ref_losses), f"wfit K{K} losses", tol=1e-5)
            out[f"losses_outer{outer}"] = torch.stack(ref_losses).numpy()
        np.savez_compressed(os.path.join(OUT, f"wfit_K{K}.npz"), **out)
        manifest["cases"].append(f"wfit_K{K}")


def main():
    os.makedirs(OUT, exist_ok=True)
    CrossEntropy2d, sig_NTM, sig_W, compute_iou = _import_reference()
    from oracle import simt_oracle as O

    manifest = {"torch": torch.__version__, "numpy": np.__version__, "cases": []}
    if len(sys.argv) == 3 and sys.argv[1] == "--only":   # add / refresh one family without touching the other fixtures
        fam = sys.argv[2]
        manifest = json.load(open(os.path.join(OUT, "MANIFEST.json")))
        manifest["cases"] = [c for c in manifest["cases"] if not c.startswith(fam + "_")]
        {"wfit": lambda: gen_wfit(sig_NTM, sig_W, O, manifest), "place": lambda: gen_placeholder(O, manifest)}[fam]()
        json.dump(manifest, open(os.path.join(OUT, "MANIFEST.json"), "w"), indent=1)
        print("wrote", fam, "cases;", len(manifest["cases"]), "golden cases in", OUT)
        return
    class_dist = np.load(os.path.join(REF, "ClassDist", "ClassDist_bapa.npy"))
    np.save(os.path.join(OUT, "ClassDist_bapa.npy"), class_dist)

    # ---- sig_NTM / sig_W (a6, a7) -------------------------------------------------
    for K in (0, 4, 15):
        torch.manual_seed(1234 + K)
        ntm = sig_NTM(19, K)
        wl = sig_W(19, K)
        with torch.no_grad():
            wl.weight.add_(0.05 * torch.randn_like(wl.weight))
        T = ntm()
        T.sum().backward()
        w_in = wl.weight.detach().clone()
        Wm = wl()
        # oracle restatement must be bit-identical to the reference modules
        T_o = O.sig_ntm_forward(ntm.NTM.detach(), class_dist, 19, K)
        W_o = O.sig_w_forward(w_in.clone())
        assert torch.equal(T_o, T.detach()), "sig_NTM restatement differs"
        assert torch.equal(W_o, Wm.detach()), "sig_W restatement differs"
        np.savez(os.path.join(OUT, f"ntm_K{K}.npz"), NTM=ntm.NTM.detach().numpy(), T=T.detach().numpy(),
                 dNTM_of_sumT=ntm.NTM.grad.numpy(), W_weight_in=w_in.numpy(), W=Wm.detach().numpy())
        manifest["cases"].append(f"ntm_K{K}")

    # ---- the head (a1-a5) ---------------------------------------------------------
    for name, B, K, h, w, H, W, coh, ign in HEAD_CASES:
        CK = 19 + K
        logits, labels = O.synth_head_inputs(B, CK, h, w, H, W, seed=zlib.crc32(name.encode()) % 10007, coherent=coh,
                                             ignore_frac=ign, class_dist=class_dist, block=8)
        torch.manual_seed(99)
        T = sig_NTM(19, K)().detach()
        out = {"logits": logits.numpy(), "labels": labels.numpy(), "T": T.numpy(), "size": np.array([H, W])}
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            loss, dl, dT = ref_head(CrossEntropy2d, logits, T, labels, (H, W), dt)
            lo, dlo, dTo = O.simt_head_fwd_bwd(logits, T, labels, (H, W), dt)
            # torch's multi-threaded CPU backward is not bit-reproducible run to run; the restatement
            # must agree to rounding (and does bit for bit on most runs)
            for x, y in ((lo, loss), (dlo, dl), (dTo, dT)):
                assert float((x.double() - y.double()).norm()) <= 1e-6 * float(y.double().norm()), name
            out[f"loss_{tag}"] = loss.numpy()
            out[f"dlogits_{tag}"] = dl.numpy()
            out[f"dT_{tag}"] = dT.numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        manifest["cases"].append(name)

    # ---- CrossEntropy2d alone, both modes (a4) --------------------------------------
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 19, 16, 24, generator=g)
    y = torch.randint(0, 19, (2, 16, 24), generator=g)
    y[torch.rand(2, 16, 24, generator=g) < 0.2] = 255
    y[0, 0, 0] = -1
    ce = {"x": x.numpy(), "y": y.numpy()}
    for mode in (True, False):
        xin = (x if mode else torch.softmax(x, 1)).clone().requires_grad_(True)
        l = CrossEntropy2d(is_softmax=mode)(xin, y)
        l.backward()
        lo = O.cross_entropy_2d(xin.detach(), y, is_softmax=mode)
        assert torch.equal(lo, l.detach())
        ce[f"loss_softmax{int(mode)}"] = l.detach().numpy()
        ce[f"grad_softmax{int(mode)}"] = xin.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "ce2d.npz"), **ce)
    manifest["cases"].append("ce2d")

    # ---- regularisers (a8-a10), executed from the cited trainV2_simt.py lines ---------
    for K in (4, 15):
        CK = 19 + K
        torch.manual_seed(4321 + K)
        ntm1, ntm2, w1, w2 = sig_NTM(19, K), sig_NTM(19, K), sig_W(19, K), sig_W(19, K)
        with torch.no_grad():
            w1.weight.add_(0.1 * torch.randn_like(w1.weight))
            w2.weight.add_(0.1 * torch.randn_like(w2.weight))
        T1, T2, W1, W2 = ntm1(), ntm2(), w1(), w2()
        for t in (T1, T2, W1, W2):
            t.retain_grad()
        loss_mse = torch.nn.MSELoss(reduction="sum")
        zeros = torch.zeros(CK, 19)
        convex = 0.0 - (loss_mse(W1.mm(T1), zeros) + loss_mse(W2.mm(T2), zeros))       # :414-415
        volume = torch.log(torch.sqrt(torch.abs(torch.linalg.det(T1.transpose(1, 0).mm(T1)))))   # :417
        volume = volume + torch.log(torch.sqrt(torch.abs(torch.linalg.det(T2.transpose(1, 0).mm(T2)))))  # :418
        # anchor, B = 1 (:354,357,375-384)
        h, w, H, W = 9, 17, 64, 128
        gg = torch.Generator().manual_seed(5)
        p1 = 3 * torch.randn(1, CK, h, w, generator=gg)
        p2 = 3 * torch.randn(1, CK, h, w, generator=gg)
        fx = 3 * torch.randn(1, 19, h, w, generator=gg)
        import torch.nn as nn
        interp = nn.Upsample(size=(H, W), mode="bilinear", align_corners=True)
        labelC = interp(torch.softmax(fx.clone(), dim=1))                               # :354
        labelC_flat = labelC.permute(0, 2, 3, 1).view(-1, 19)                           # :357
        anchor = 0.0
        a_idx, exists = [], []
        for pr, T in ((p1, T1), (p2, T2)):
            pu = interp(pr)
            flat = pu.clone().permute(0, 2, 3, 1).view(-1, CK).detach()                 # :375
            ai = torch.argmax(flat, dim=0)                                              # :376
            ex = torch.unique(torch.argmax(flat, dim=1))                                # :377
            anc = labelC_flat[ai]                                                       # :378
            anchor = anchor + loss_mse(T[ex], anc[ex])                                  # :379
            a_idx.append(ai.numpy()); exists.append(ex.numpy())
        total = 0.1 * convex + 1.0 * volume + 1.0 * anchor                              # :424 weights of sh_simt.sh:16
        total.backward()
        # oracle restatement check
        co = O.convex_loss([W1.detach(), W2.detach()], [T1.detach(), T2.detach()])
        vo = O.volume_loss([T1.detach(), T2.detach()])
        ao = O.anchor_loss([interp(p1), interp(p2)], [T1.detach(), T2.detach()], O.label_c_flat(fx, (H, W)))
        assert torch.equal(co, convex.detach()) and torch.equal(vo, volume.detach()) and torch.equal(ao, anchor.detach())
        np.savez_compressed(
            os.path.join(OUT, f"reg_K{K}.npz"),
            T1=T1.detach().numpy(), T2=T2.detach().numpy(), W1=W1.detach().numpy(), W2=W2.detach().numpy(),
            convex=convex.detach().numpy(), volume=volume.detach().numpy(), anchor=anchor.detach().numpy(),
            dT1=T1.grad.numpy(), dT2=T2.grad.numpy(), dW1=W1.grad.numpy(), dW2=W2.grad.numpy(),
            p1=p1.numpy(), p2=p2.numpy(), fixed=fx.numpy(), size=np.array([H, W]),
            anchor_idx1=a_idx[0], anchor_idx2=a_idx[1], exist1=exists[0], exist2=exists[1])
        manifest["cases"].append(f"reg_K{K}")

    gen_wfit(sig_NTM, sig_W, O, manifest)
    gen_placeholder(O, manifest)

    # ---- pseudo labels + class-posterior relabel (section 8(f) row 2), executed from trainV2_simt.py:354-365,387-393 ----
    import torch.nn as nn
    gg = torch.Generator().manual_seed(23)
    Bp, Cp, CKp, hp, wp, Hp, Wp = 2, 19, 23, 9, 17, 64, 128
    output2 = 2.0 * torch.randn(Bp, Cp, hp, wp, generator=gg)
    pred2_lo = 2.0 * torch.randn(Bp, CKp, hp, wp, generator=gg)
    interp_target = nn.Upsample(size=(Hp, Wp), mode="bilinear", align_corners=True)
    num_classes = 19
    labelC = interp_target(torch.softmax(output2.clone(), dim=1))                                           # :354
    labelC_max = torch.max(labelC, 1)                                                                        # :355
    labelC_argmax = torch.argmax(labelC, dim=1).float()                                                      # :356
    labelC = torch.where(labelC_max[0] > 0.8, labelC_argmax, 255. * torch.ones_like(labelC_argmax))          # :359
    labelC = torch.where(labelC_max[0] < 0.2, num_classes * torch.ones_like(labelC_argmax), labelC)          # :361
    Conf = torch.from_numpy(labelC.detach().clone().cpu().numpy()).long()                                    # :362
    pred2 = interp_target(pred2_lo)                                                                          # :372
    pseudo = torch.argmax(pred2.clone(), dim=1).detach()                                                     # :387
    ones = torch.ones_like(Conf); zeros = torch.zeros_like(Conf)
    mask = torch.where(Conf == num_classes * ones, ones, zeros)                                              # :390
    pseudo1 = mask * pseudo                                                                                  # :391
    pseudo1 = torch.where(pseudo1 >= num_classes * ones, pseudo1, 255 * ones)                                # :392
    Conf = torch.where(Conf == num_classes * ones, pseudo1, Conf)                                            # :393
    assert torch.equal(Conf, O.pseudo_labels(output2, pred2, (Hp, Wp), num_classes, 0.8, 0.2))
    np.savez_compressed(os.path.join(OUT, "pseudo_K4.npz"), output2=output2.numpy(), pred2_lo=pred2_lo.numpy(),
                        size=np.array([Hp, Wp]), conf=Conf.numpy().astype(np.uint8))
    manifest["cases"].append("pseudo_K4")

    # ---- histograms (a12-a16): compute_iou's functions imported unmodified -----------------
    info = json.load(open(os.path.join(REF, "dataset", "cityscapes_list", "info.json")))
    mapping = np.array(info["label2train"], dtype=np.int64)
    assert mapping.tolist() == O.CITYSCAPES_LABEL2TRAIN and int(info["classes"]) == 19
    hist = np.zeros((19, 19))
    pairs = []
    for i in range(3):
        gt, pr = O.synth_eval_pair(96, 160, seed=100 + i, coherent=(i != 1), block=16)
        lab = compute_iou.label_mapping(gt, mapping)
        assert np.array_equal(lab, O.label_mapping(gt, mapping))
        hi = compute_iou.fast_hist(lab.flatten(), pr.flatten(), 19)
        assert np.array_equal(hi, O.fast_hist(lab.flatten(), pr.flatten(), 19))
        hist += hi
        pairs.append((gt, pr))
    iu = compute_iou.per_class_iu(hist)
    miou = round(np.nanmean(iu) * 100, 2)
    assert miou == O.miou_percent(hist)
    gt0, pr0 = pairs[0]
    ka = (gt0.flatten() >= 0) & (gt0.flatten() < 34)                                         # compute_ConfusionMatrix.py:55
    rect = np.bincount(19 * gt0.flatten()[ka].astype(int) + pr0.flatten()[ka], minlength=34 * 19).reshape(34, 19)  # :56
    kb = (pr0.flatten() >= 0) & (pr0.flatten() < 19)                                         # compute_ClassDistribution.py:53
    cd = np.bincount(pr0.flatten()[kb], minlength=19)                                         # :54
    assert np.array_equal(rect, O.fast_hist_rect(gt0.flatten(), pr0.flatten(), 34, 19))
    assert np.array_equal(cd, O.class_hist(pr0.flatten(), 19))
    np.savez_compressed(os.path.join(OUT, "hist.npz"),
                        gt=np.stack([p[0] for p in pairs]), pred=np.stack([p[1] for p in pairs]),
                        mapping=mapping, hist19=hist.astype(np.int64), iu=iu, miou=np.float64(miou),
                        rect34x19=rect.astype(np.int64), class19=cd.astype(np.int64))
    manifest["cases"].append("hist")

    json.dump(manifest, open(os.path.join(OUT, "MANIFEST.json"), "w"), indent=1)
    print("wrote", len(manifest["cases"]), "golden cases to", OUT)


if __name__ == "__main__":
    main()
