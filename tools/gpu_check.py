"""GPU diagnostic: parity reports + kernel timings for a list of configs. Writes gpurun_out/gpu_check.json."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers as h

def main():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    print(torch.cuda.get_device_name(0), flush=True)
    cases = [
        ("tiny_deg0", dict(P=64, res=32, deg=0, sigma=0.08)),
        ("small_deg3", dict(P=300, res=64, deg=3, sigma=0.05, elev=10, azim=30)),
        ("odd_size", dict(P=500, res=0, width=100, height=70, deg=1, sigma=0.04, elev=-20, azim=200)),
        ("cfg1_5k_256", dict(P=5000, res=256, deg=0, opacity="init", anisotropic=False)),
        ("big_gauss", dict(P=500, res=80, deg=2, sigma=0.2, elev=25, azim=-100)),
        ("mid_20k_400", dict(P=20000, res=400, deg=3)),
    ]
    cases += [
        ("big_tiles_radix", dict(P=20000, res=32, deg=0, sigma=0.02)),            # ~5k instances per tile: big-list radix path
        ("huge_tiles_bitonic", dict(P=60000, res=32, deg=0, sigma=0.01)),         # > 11264 per tile: in-place global network
        ("planar_equal_depth", dict(P=3000, res=64, deg=1, sigma=0.03, planar=True)),  # all depths equal: degenerate-run fallback
    ]
    if "--tiny" in sys.argv:
        cases = cases[:3]
    if "--full" in sys.argv:
        cases.append(("cfg2_100k_800", dict(P=100000, res=800, deg=3)))
    allrep = {}
    for name, kw in cases:
        planar = kw.pop("planar", False)
        s, i = h.make_case(**kw)
        if planar:      # squash the cloud onto the plane z = 0 seen head-on: every Gaussian has the same view depth
            i["means3D"] = i["means3D"].copy(); i["means3D"][:, 2] = 0.0
        g = h.upstream_grads(s["image_height"], s["image_width"])
        t0 = time.time(); ref = h.run_oracle(s, i, g); t1 = time.time()
        try:
            cu = h.run_cuda(s, i, g)
            torch.cuda.synchronize()
            ok, rep = h.compare(cu, ref)
        except Exception as e:
            ok, rep = False, {"exception": repr(e)}
        rep["n_inst_oracle"] = ref["n_inst"]; rep["oracle_s"] = round(t1 - t0, 3); rep["ok"] = ok
        allrep[name] = rep
        print(name, "OK" if ok else "FAIL", json.dumps({k: (round(v, 8) if isinstance(v, float) else v) for k, v in rep.items()}), flush=True)
    json.dump(allrep, open(os.path.join(ROOT, "gpurun_out", "gpu_check.json"), "w"), indent=1)

if __name__ == "__main__":
    main()
