"""The two cases compute-sanitizer runs (tools/r2_sanitize.sh): forward + backward through the public API, checked against
the oracle so that a sanitizer-clean run is also a correct one."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as h

for name, kw in (("cfg1_5k_256_init", dict(P=5000, res=256, deg=0, opacity="init", anisotropic=False)),
                 ("small_deg3", dict(P=300, res=64, deg=3, sigma=0.05, elev=10, azim=30)),
                 ("big_tiles", dict(P=20000, res=32, deg=0, sigma=0.02))):
    s, i = h.make_case(**kw)
    g = h.upstream_grads(s["image_height"], s["image_width"])
    ok, rep = h.compare(h.run_cuda(s, i, g), h.run_oracle(s, i, g))
    print(name, "ok" if ok else "FAIL %r" % rep, flush=True)
    assert ok
print("cases ok")
