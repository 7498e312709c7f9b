"""Config 2 / 4 of BASELINE.json at the size it names (~1e5 - 2e5 DoF): the Miehe shear test (stress split) on a
uniformly refined slit square through the C++ command line, which preconditions with the 2-D multigrid from
64 x 64 cells on (pf_set_preconditioner kind 3).  Prints wall time, Newton and GMRES iteration counts.
usage: python tools/miehe_scale.py [--refine 7] [--steps 3]        (refine 7 = 256 x 256 cells, 198 531 DoF)"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from prm_from_golden import write_prm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--refine", type=int, default=7)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--case", default="shear", choices=["shear", "tension"],
                    help="shear: parameters_miehe_shear (stress split, BASELINE config 4 physics); tension: "
                         "parameters_miehe_tension_adaptive on the uniform mesh (BASELINE config 2)")
    ap.add_argument("--exe", default=os.path.join(ROOT, "cracks_b200", "cracks_b200_run"))
    args = ap.parse_args()
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "miehe_shear_2.json" if args.case == "shear"
                                    else "miehe_tension_adaptive_1.json")))
    with tempfile.TemporaryDirectory() as tmp:
        prm = write_prm(os.path.join(tmp, "m.prm"), g["prm"], 2, os.path.join(tmp, "out"), Max_No_of_timesteps=args.steps - 1,
                        exact={"Global pre-refinement steps": args.refine, "Adaptive refinement cycles": 0})
        t0 = time.time()
        r = subprocess.run([args.exe, prm, "--no-output"], capture_output=True, text=True)
        wall = time.time() - t0
        if r.returncode != 0:
            print(r.stdout[-2000:], r.stderr[-2000:])
            sys.exit(r.returncode)
        its = [tuple(map(int, m)) for m in re.findall(r"Newton iterations: (\d+) total linear iterations: (\d+)", r.stdout)]
        dofs = re.search(r"DoFs: .* = (\d+)", r.stdout).group(1)
        rows = [l.split() for l in open(os.path.join(tmp, "out", "statistics")) if not l.startswith("#")]
        print(json.dumps({"case": "miehe " + args.case, "n_dofs": int(dofs), "time_steps": len(rows), "newton_its": sum(a for a, _ in its),
                          "linear_its": sum(b for _, b in its), "wall_s": round(wall, 2),
                          "last_row": {"bulk": float(rows[-1][4]), "crack": float(rows[-1][5]), "load": float(rows[-1][6])}}))


if __name__ == "__main__":
    main()
