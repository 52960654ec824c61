"""Golden vectors for the reference's host-side curve utilities, generated from the LIVE reference in the build container:
block_stitch_sfc (/root/reference/src/curves/space_filling_curves.py:513-591) for 4 curves x several grids (incl. non-square)
and refine_curve_to_hamiltonian (:446-455) / find_hamiltonian_path (:273-443). Writes tests/golden/host_curves.json:
sha256[:16] of the int64 flat index array (i * height + j) plus block lengths / counts. Run: python tests/golden/make_host_curve_golden.py"""
import hashlib
import json
import os
import signal
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import  # noqa: E402


def h(curve, height):
    return hashlib.sha256(np.array([int(i) * height + int(j) for i, j in curve], "<i8").tobytes()).hexdigest()[:16]


class Timeout(Exception):
    pass


def _alarm(*_):
    raise Timeout()


def main():
    ref = ref_import.load_reference()["curves"]
    curves = {"hilbert": ref.hilbert_curve, "z": ref.z_curve, "peano": ref.peano_curve, "moore": ref.moore_curve}
    out = {"block_stitch": {}, "refine": {}, "hamiltonian": {}}
    for name, fn in curves.items():
        for (w, hh) in [(7, 7), (12, 12), (14, 14), (24, 24), (6, 10), (20, 9), (27, 27), (32, 32)]:
            curve, blocked = ref.block_stitch_sfc(fn, w, hh)
            out["block_stitch"][f"{name}_{w}x{hh}"] = {"hash": h(curve, hh), "n": len(curve), "blocks": [len(b) for b in blocked]}
    signal.signal(signal.SIGALRM, _alarm)
    for name, fn in curves.items():
        for (w, hh) in [(7, 7), (12, 12), (14, 14), (24, 24), (6, 10)]:
            key = f"{name}_{w}x{hh}"
            signal.alarm(20)
            try:
                t0 = time.time()
                ham = ref.refine_curve_to_hamiltonian(ref.embed_and_prune_sfc(fn, w, hh), w, hh)
                out["refine"][key] = {"hash": h(ham, hh) if ham else None, "n": len(ham) if ham else 0, "ref_seconds": round(time.time() - t0, 3)}
            except Timeout:
                out["refine"][key] = {"timeout_s": 20}
            except RecursionError:
                out["refine"][key] = {"recursion_error": True}
            finally:
                signal.alarm(0)
    for (w, hh, diag) in [(4, 4, False), (5, 7, False), (8, 8, False), (6, 6, True)]:
        signal.alarm(20)
        try:
            p = ref.find_hamiltonian_path(w, hh, diag=diag)
            out["hamiltonian"][f"{w}x{hh}_diag{int(diag)}"] = {"hash": h(p, hh) if p else None, "n": len(p) if p else 0}
        except Timeout:
            out["hamiltonian"][f"{w}x{hh}_diag{int(diag)}"] = {"timeout_s": 20}
        finally:
            signal.alarm(0)
    with open(os.path.join(HERE, "host_curves.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out["refine"], indent=1))
    print({k: v["hash"] for k, v in list(out["block_stitch"].items())[:6]})
    print(out["hamiltonian"])


if __name__ == "__main__":
    main()
