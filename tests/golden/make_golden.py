#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED compiled reference.

Run in the build container only (needs /root/reference -> oracle/_ref, see
oracle/build.py):   python tests/golden/make_golden.py

Each file holds the inputs (float64 frames, masks, positions, parameters) and
what the reference's own UMPAModel*.match() / hooks returned for them, so the
C oracle (CPU tests) and the CUDA path (GPU tests) can be pinned without the
reference being present.  The reference's test-suite has no golden vectors of
its own; these are the substitute.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref as oref          # noqa: E402
from umpa_b200 import synth             # noqa: E402

R = oref.load()
if R is None:
    sys.exit("compiled reference not available (oracle/build.py needs /root/reference)")

CLS = {"NoDF": R.UMPAModelNoDF, "DF": R.UMPAModelDF, "DFKernel": R.UMPAModelDFKernel}
KEYS = ("f", "T", "dx", "dy", "df", "err", "debug_Ncalls", "debug_d", "debug_a")
ONLY = set(sys.argv[1:])            # optional: regenerate just the named cases


def run_case(name, kind, sam, ref, mask=None, pos=None, Nw=2, max_shift=4, abc=None,
             assign=None, subpx=None, step=None, ROI=None, dxdy=None):
    if ONLY and name not in ONLY:
        return
    sam_l = [np.ascontiguousarray(s) for s in sam]
    ref_l = [np.ascontiguousarray(r) for r in ref]
    mask_l = None if mask is None else [np.ascontiguousarray(m) for m in mask]
    pos_l = None if pos is None else [np.array(p) for p in pos]
    m = CLS[kind](sam_l, ref_l, mask_list=mask_l, pos_list=pos_l, window_size=Nw, max_shift=max_shift)
    if assign is not None:
        m.assign_coordinates = assign
    if subpx is not None:
        m.sub_pixel_mode = subpx
    kw = dict(num_threads=4, quiet=True)
    if step is not None:
        kw["step"] = step
    if ROI is not None:
        kw["ROI"] = ROI
    if dxdy is not None:
        kw["dxdy"] = dxdy
    if kind == "DFKernel":
        if abc is None:
            abc = synth.blur_abc(*m.sh)
        kw["abc"] = abc
    res = m.match(**kw)
    out = {"kind": kind, "Nw": Nw, "max_shift": max_shift, "padding": m.padding,
           "extent": np.array(m.extent), "ROI_after": np.array(m.ROI), "sh_after": np.array(m.sh),
           "window": np.array(m.window)}
    # frames may be ragged (sample stepping): store one entry per frame
    out["Na"] = len(sam_l)
    for k, (s, r) in enumerate(zip(sam_l, ref_l)):
        out[f"sam{k}"], out[f"ref{k}"] = s, r
        if mask_l is not None:
            out[f"mask{k}"] = mask_l[k]
    if pos_l is not None:
        out["pos"] = np.array(pos_l)
    for k_, v in (("abc", abc), ("assign", assign), ("subpx", subpx), ("step", step),
                  ("ROI", None if ROI is None else np.array(ROI)), ("dxdy", dxdy)):
        if v is not None:
            out[k_] = v
    for k_ in KEYS:
        if k_ in res:
            out["out_" + k_] = res[k_]
    # a few single-pixel probes through the reference's cost()/min() hooks
    p = m.padding
    probes = []
    rng = np.random.default_rng(7)
    H0, W0 = sam_l[0].shape
    for _ in range(6):
        i = int(rng.integers(p, H0 - p)); j = int(rng.integers(p, W0 - p))
        si = int(rng.integers(-max_shift + 1, max_shift)); sj = int(rng.integers(-max_shift + 1, max_shift))
        if kind == "DFKernel":
            c = m.cost(i, j, si, sj, .5, .1, .4)
        else:
            c = m.cost(i, j, si, sj)
        probes.append([i, j, si, sj] + list(c) + [0.] * (3 - len(c)))
    out["cost_probes"] = np.array(probes)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    ok = res["err"] == 1
    print(f"{name:22s} {kind:9s} out {res['f'].shape} ok {ok.mean():.3f} "
          f"calls {res['debug_Ncalls'][ok].mean() if ok.any() else 0:.2f}")


def load_reference_align():
    """The reference's own UMPA/align.py, imported from its file with the modules it does not need for
    correct_bad_pixels stubbed (matplotlib is absent here; `import UMPA` would pull the whole package)."""
    import importlib.util
    import types
    for name in ("matplotlib", "matplotlib.pyplot", "UMPA"):
        sys.modules.setdefault(name, types.ModuleType(name))
    spec = importlib.util.spec_from_file_location("ref_align", "/root/reference/UMPA/align.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_post():
    """correct_bad_pixels (align.py:661-732) on displacement-like maps with outliers on edges and corners."""
    A = load_reference_align()
    rng = np.random.default_rng(21)
    out = {}
    img = rng.normal(0., .8, (40, 44))
    img[rng.random(img.shape) < .03] = 7.5
    img[rng.random(img.shape) < .03] = -9.
    img[0, 0], img[0, 43], img[39, 0], img[39, 43], img[0, 10], img[39, 11], img[12, 0], img[13, 43] = 8, -8, 9, -9, 6, -6, 5, -5
    img[20:23, 20:23] = 11.                       # a cluster: neighbours are bad too
    stack = rng.normal(0., 1., (3, 20, 22))
    stack[rng.random(stack.shape) < .05] = 6.
    out["img"], out["stack"] = img, stack
    out["img_th3_it1"] = A.correct_bad_pixels(img, 3)
    out["img_th3_it2"] = A.correct_bad_pixels(img, 3, iterations=2)
    out["img_th3_it3"] = A.correct_bad_pixels(img, 3, iterations=3)
    out["img_auto"] = A.correct_bad_pixels(img)
    out["img_auto_p5"] = A.correct_bad_pixels(img, p=5.)
    out["stack_th4"] = A.correct_bad_pixels(stack, 4)
    out["clean"] = A.correct_bad_pixels(np.clip(img, -2, 2), 3)
    np.savez_compressed(os.path.join(HERE, "post_badpix.npz"), **out)
    print("post_badpix written")


def main():
    clean = synth.speckle_stack(5, 40, 44, seed=11, max_shift=4, dark_field=False)
    clean_df = synth.speckle_stack(5, 40, 44, seed=12, max_shift=4, dark_field=True)
    noisy = synth.speckle_stack(4, 40, 44, seed=13, max_shift=4, dark_field=True, noise=.2, amplitude=2.5)
    lowc = synth.speckle_stack(5, 40, 44, seed=14, max_shift=4, dark_field=True, contrast=.15)
    big = synth.speckle_stack(3, 50, 54, seed=15, max_shift=4, dark_field=True)
    wide = synth.speckle_stack(4, 44, 48, seed=16, max_shift=6, dark_field=True, amplitude=3.2)

    run_case("nodf_clean", "NoDF", clean["sam"], clean["ref"])
    run_case("df_clean", "DF", clean_df["sam"], clean_df["ref"])
    run_case("nodf_noisy", "NoDF", noisy["sam"], noisy["ref"])
    run_case("df_noisy", "DF", noisy["sam"], noisy["ref"])
    run_case("df_lowcontrast", "DF", lowc["sam"], lowc["ref"])
    run_case("df_nw3_ms6", "DF", wide["sam"], wide["ref"], Nw=3, max_shift=6)
    run_case("nodf_nw1", "NoDF", clean["sam"], clean["ref"], Nw=1, max_shift=3)
    run_case("dfk_clean", "DFKernel", big["sam"], big["ref"], Nw=2, max_shift=3)

    # DFKernel at the shape class of BASELINE config 3 (Nw=3, max_shift=5), a strided ROI, and a wide
    # blur whose 17x17 truncation matters (a, c ~ 0.03-0.08)
    kbig = synth.speckle_stack(6, 72, 76, seed=17, max_shift=5, dark_field=True)
    run_case("dfk_nw3_ms5", "DFKernel", kbig["sam"], kbig["ref"], Nw=3, max_shift=5)
    run_case("dfk_roi_step", "DFKernel", kbig["sam"], kbig["ref"], Nw=2, max_shift=4,
             ROI=((1, 40, 2), (3, 44, 3)), abc=synth.blur_abc(20, 14))
    wide_abc = synth.blur_abc(44, 48)
    wide_abc[..., 0] *= .12
    wide_abc[..., 2] *= .08
    wide_abc[..., 1] *= .2
    run_case("dfk_wide_blur", "DFKernel", kbig["sam"], kbig["ref"], Nw=1, max_shift=5, abc=wide_abc)
    run_case("dfk_assign_ref", "DFKernel", kbig["sam"], kbig["ref"], Nw=2, max_shift=4, assign="ref")
    run_case("dfk_assign_ref_step", "DFKernel", kbig["sam"], kbig["ref"], Nw=2, max_shift=5, assign="ref",
             ROI=((0, 40, 3), (2, 44, 2)), abc=synth.blur_abc(14, 21))

    # masks: smooth positive weights with a dead block and a few dead pixels
    rng = np.random.default_rng(5)
    mask = .5 + .5 * rng.random(clean_df["sam"].shape)
    mask[:, 10:14, 20:26] = 0.
    mask[rng.random(mask.shape) < .02] = 0.
    run_case("nodf_masked", "NoDF", clean["sam"], clean["ref"], mask=mask)
    run_case("df_masked", "DF", clean_df["sam"], clean_df["ref"], mask=mask)
    mask_b = .5 + .5 * rng.random(big["sam"].shape)
    run_case("dfk_masked", "DFKernel", big["sam"], big["ref"], mask=mask_b, Nw=1, max_shift=3)

    # sparse masks (what real masks look like: ones, a dead block, a dead pixel, a down-weighted corner):
    # most pixels have no mask value != 1 within reach
    sp = synth.speckle_stack(5, 72, 76, seed=18, max_shift=4, dark_field=True)
    msp = np.ones(sp["sam"].shape)
    msp[1, 30:33, 40:43] = 0.
    msp[3, 10, 12] = 0.
    msp[:, 58:, :9] = .5
    run_case("df_masked_sparse", "DF", sp["sam"], sp["ref"], mask=msp)
    run_case("nodf_masked_sparse", "NoDF", sp["sam"], sp["ref"], mask=msp)
    run_case("df_masked_sparse_ref", "DF", sp["sam"], sp["ref"], mask=msp, assign="ref")

    # DFKernel with a sparse mask / with sample stepping (the blur widens the reach by 8 pixels)
    run_case("dfk_masked_sparse", "DFKernel", sp["sam"], sp["ref"], mask=msp, Nw=1, max_shift=4)

    # sample stepping: ragged frames at integer offsets (model.pyx:265-283)
    pos = [(0, 0), (3, 0), (0, 5), (2, 2), (5, 4)]
    shapes = [(40, 44), (38, 44), (40, 40), (36, 42), (35, 40)]
    sam_p = [clean_df["sam"][k][:h, :w].copy() for k, (h, w) in enumerate(shapes)]
    ref_p = [clean_df["ref"][k][:h, :w].copy() for k, (h, w) in enumerate(shapes)]
    run_case("df_positions", "DF", sam_p, ref_p, pos=pos)
    run_case("nodf_positions", "NoDF", sam_p, ref_p, pos=pos)

    # larger sample-stepping set: 6 ragged frames at offsets up to 14 px (interior reached by all frames,
    # rims reached by some, corners by none)
    stp = synth.speckle_stack(6, 96, 100, seed=19, max_shift=4, dark_field=True)
    pos6 = [(0, 0), (9, 0), (0, 14), (6, 7), (12, 3), (3, 11)]
    shp6 = [(84, 86), (80, 90), (90, 80), (84, 84), (78, 88), (88, 82)]
    sam6 = [stp["sam"][k][py:py + h, px:px + w].copy() for k, ((py, px), (h, w)) in enumerate(zip(pos6, shp6))]
    ref6 = [stp["ref"][k][py:py + h, px:px + w].copy() for k, ((py, px), (h, w)) in enumerate(zip(pos6, shp6))]
    run_case("df_positions_big", "DF", sam6, ref6, pos=pos6)
    run_case("nodf_positions_big", "NoDF", sam6, ref6, pos=pos6, assign="ref")
    run_case("dfk_positions_big", "DFKernel", sam6, ref6, pos=pos6, Nw=1, max_shift=4)

    # options
    run_case("df_assign_ref", "DF", clean_df["sam"], clean_df["ref"], assign="ref")
    run_case("df_subpx0", "DF", clean_df["sam"], clean_df["ref"], subpx=0)
    run_case("df_subpx1", "DF", clean_df["sam"], clean_df["ref"], subpx=1)
    run_case("df_step3", "DF", clean_df["sam"], clean_df["ref"], step=3)
    run_case("df_roi", "DF", clean_df["sam"], clean_df["ref"], ROI=((2, 20, 2), (1, 25, 3)))
    run_case("df_dxdy", "DF", clean_df["sam"], clean_df["ref"], dxdy=(1., -1.))

    if not ONLY or "post_badpix" in ONLY:
        make_post()
    if ONLY and "hooks" not in ONLY:
        return
    # module-level hooks (model.pyx:31-114; note spm -> spmin_quad, spmq -> spmin)
    blocks = []
    for n in range(8):
        g = np.random.default_rng(100 + n)
        yy, xx = np.mgrid[-1:3, -1:3].astype(float)
        cy, cx = g.uniform(0, 1, 2)
        a = .3 + (yy - cy) ** 2 + .8 * (xx - cx) ** 2 + .3 * (yy - cy) * (xx - cx) + .02 * g.random((4, 4))
        pq, vq = R.spm(np.ascontiguousarray(a))
        ps, vs = R.spmq(np.ascontiguousarray(a))
        blocks.append(np.concatenate([a.ravel(), pq, [vq], ps, [vs]]))
    kern = R.test_CostArgsDFKernel(0, 0, .45, .12, .6)
    np.savez_compressed(os.path.join(HERE, "hooks.npz"), blocks=np.array(blocks), kernel=kern,
                        kernel_abc=np.array([.45, .12, .6]))
    print("hooks written")


if __name__ == "__main__":
    main()
