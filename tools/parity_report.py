"""Parity report on a GPU: every suite case + the BASELINE shapes against BOTH builds of the CPU oracle.
  * float64 oracle: tests/helpers.compare (decision-ambiguity flags, tie slack);
  * float32 oracle: raw deviations with NO exemption and with depth-key ties exempted only.
Writes gpurun_out/parity_report.json (copied to profiles/ when it backs a tolerance in tests/helpers.py)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import helpers as h

CASES = {
    "tiny_deg0": dict(P=64, res=32, deg=0, sigma=0.08),
    "small_deg3": dict(P=300, res=64, deg=3, sigma=0.05, elev=10, azim=30),
    "odd_size_deg1": dict(P=500, res=0, width=100, height=70, deg=1, sigma=0.04, elev=-20, azim=200),
    "cfg1_5k_256_init": dict(P=5000, res=256, deg=0, opacity="init", anisotropic=False),
    "big_gaussians_deg2": dict(P=500, res=80, deg=2, sigma=0.2, elev=25, azim=-100),
    "mid_20k_400_deg3": dict(P=20000, res=400, deg=3),
    "scale_modifier": dict(P=400, res=64, deg=1, sigma=0.04, scale_modifier=1.6, bg=(0.2, 0.7, 0.1)),
    "big_tiles": dict(P=20000, res=32, deg=0, sigma=0.02),
    "huge_tiles": dict(P=60000, res=32, deg=0, sigma=0.01),
    "cfg2_100k_800": dict(P=100000, res=800, deg=3),
    "cfg2_100k_800_init": dict(P=100000, res=800, deg=3, opacity="init"),
    "cfg3_view_500k_512": dict(P=500000, res=512, deg=3, sigma=0.0075, elev=-12, azim=75),
}
if "--big" in sys.argv:
    CASES["cfg5_2M_1600"] = dict(P=2000000, res=1600, deg=3, sigma=0.004, elev=10, azim=30)


def raw_dev(cu, ref, exempt_px, exempt_g):
    out = {}
    for name in ("color", "depth", "alpha"):
        d = np.abs(cu[name].astype(np.float64) - ref[name].astype(np.float64))
        if name == "depth":
            d = d / max(1.0, float(np.abs(ref[name]).max()))
        m = np.broadcast_to(exempt_px[None], d.shape)
        dd = d[~m]
        out[name] = dict(max=float(dd.max()) if dd.size else 0.0, n_gt_1e4=int((dd > 1e-4).sum()), n_gt_1e3=int((dd > 1e-3).sum()),
                         n_gt_1e2=int((dd > 1e-2).sum()), n=int(dd.size))
    out["radii_mismatch"] = int(((cu["radii"] != ref["radii"]) & ~exempt_g).sum())
    floor = 1e-3 * max(float(np.abs(v).max()) if v.size else 0.0 for v in ref["grads"].values())
    for k, gr in ref["grads"].items():
        if k not in cu["grads"] or gr.size == 0:
            continue
        gc = cu["grads"][k].astype(np.float64).reshape(gr.shape)
        scale = max(float(np.abs(gr).max()), floor) or 1.0
        e = (np.abs(gc - gr.astype(np.float64)) - h.GRAD_RTOL * np.abs(gr)).reshape(gr.shape[0], -1).max(axis=1) / scale
        e = e[~exempt_g]
        out["grad_" + k] = dict(max=float(e.max()) if e.size else 0.0, p999=float(np.percentile(e, 99.9)) if e.size else 0.0,
                                n_gt_1e4=int((e > 1e-4).sum()), n=int(e.size))
    return out


def main():
    res = {}
    for name, kw in CASES.items():
        t0 = time.time()
        s, i = h.make_case(**kw)
        depth = not name.startswith("cfg3") and not name.startswith("cfg5")
        g = h.upstream_grads(s["image_height"], s["image_width"], depth=depth)
        cu = h.run_cuda(s, i, g)
        r64 = h.run_oracle(s, i, g)
        r32 = h.run_oracle(s, i, g, dtype=np.float32)
        ok64, rep64 = h.compare(cu, r64, max_ambig_frac=1.0)
        none_px = np.zeros_like(r32["ambig_px"], bool); none_g = np.zeros_like(r32["ambig_g"], bool)
        tie_px = ((r32["ambig_px"] | r64["ambig_px"]) & 4) != 0
        tie_g = ((r32["ambig_g"] | r64["ambig_g"]) & (2 | 16)) != 0
        rad_g = ((r32["ambig_g"] | r64["ambig_g"]) & 4) != 0
        res[name] = dict(vs_f64=dict(ok=ok64, **rep64),
                         vs_f32_no_exemption=raw_dev(cu, r32, none_px, none_g),
                         vs_f32_ties_exempt=raw_dev(cu, r32, tie_px, tie_g | rad_g),
                         tie_px_frac=float(tie_px.mean()), tie_g_frac=float(tie_g.mean()), n_inst=int(r64["n_inst"]),
                         seconds=time.time() - t0)
        print(name, "f64 ok" if ok64 else "f64 FAIL", json.dumps(res[name]["vs_f32_ties_exempt"])[:600], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "parity_report.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
