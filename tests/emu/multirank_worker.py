"""One rank of a multi-process run of the EMULATED library (tests/test_multirank_emulation_cpu.py):
python multirank_worker.py <rank> <nranks> <id_file> <out_file> <nx> <ny> <nz> <time_steps>"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    rank, nranks = int(sys.argv[1]), int(sys.argv[2])
    id_file, out_file = sys.argv[3], sys.argv[4]
    n = tuple(int(v) for v in sys.argv[5:8])
    steps = int(sys.argv[8])
    from cracks_b200 import api
    import cracks_b200 as pf
    api.library_path = lambda: os.path.join(ROOT, "tests", "emu", "libcracks_b200_emu.so")
    api._LIB = None
    nccl_id = None
    if nranks > 1:
        if rank == 0:
            nccl_id = pf.PhaseFieldContext.nccl_unique_id()
            with open(id_file + ".tmp", "wb") as f:
                f.write(nccl_id)
            os.rename(id_file + ".tmp", id_file)
        else:
            for _ in range(3000):
                if os.path.exists(id_file):
                    break
                time.sleep(0.01)
            nccl_id = open(id_file, "rb").read()
    mesh = pf.Mesh()
    mesh.dim = 3
    for d in range(3):
        mesh.n[d], mesh.h[d], mesh.origin[d] = n[d], 20.0 / n[d], -10.0
    h = api.mesh_diameter(mesh)
    mu = 1.0 / 2.4
    params = pf.Params(0.4 * mu / 0.6, mu, 1.0, 1e-8 * h, 2.0 * h, 0.0)
    ctx = pf.PhaseFieldContext(mesh, params, device=0, rank=rank, nranks=nranks, nccl_id=nccl_id)
    drv = pf.SneddonDriver(ctx, pressure=lambda t: 1e-3, max_no_timesteps=steps - 1, newton_lower_bound=1e-7, max_newton=50,
                           max_line_search=10, gmres_max_it=200)
    stats = drv.run(h)
    levels = api.mg_hierarchy(mesh, rank, nranks)
    if rank == 0:
        json.dump(dict(statistics=stats, newton_its=drv.newton_its, linear_its=drv.lin_its,
                       levels=[(L["n"], L["replicated"], L["mode_below"]) for L in levels]), open(out_file, "w"))
    ctx.close()


if __name__ == "__main__":
    main()
