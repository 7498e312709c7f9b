"""One rank of a multi-process run of the EMULATED library on a FOREST mesh (pf_create_forest_distributed:
replicated vectors, partitioned cells, all-reduce through tests/emu/fake_nccl):
python multirank_forest_worker.py <rank> <nranks> <id_prefix> <out_file> <case> [max_steps]
case = hetero (tests/golden/hetero_3d_1.json, KAT-5) | shear (tests/golden/miehe_shear_1.json, adaptive)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, nranks = int(sys.argv[1]), int(sys.argv[2])
    id_prefix, out_file, case = sys.argv[3], sys.argv[4], sys.argv[5]
    max_steps = int(sys.argv[6]) if len(sys.argv) > 6 else None
    from cracks_b200 import api
    import cracks_b200 as pf
    api.library_path = lambda: os.path.join(ROOT, "tests", "emu", "libcracks_b200_emu.so")
    api._LIB = None
    counter = [0]

    def fresh_id():
        """a new ncclUniqueId per context: rank 0 writes file k, the others wait for it"""
        path = "%s_%d" % (id_prefix, counter[0])
        counter[0] += 1
        if rank == 0:
            nccl_id = pf.PhaseFieldContext.nccl_unique_id()
            with open(path + ".tmp", "wb") as f:
                f.write(nccl_id)
            os.rename(path + ".tmp", path)
            return nccl_id
        for _ in range(60000):
            if os.path.exists(path):
                break
            time.sleep(0.01)
        return open(path, "rb").read()

    dist = (rank, nranks, fresh_id)
    from forest_cases import hetero_driver, miehe_forest_driver
    if case == "hetero":
        drv, g = hetero_driver(pf, dist)
        stats = drv.run()
    else:
        g = json.load(open(os.path.join(ROOT, "tests", "golden", "miehe_shear_1.json")))
        drv = miehe_forest_driver(pf, g, max_steps=max_steps, dist=dist)
        stats = drv.run()
    if rank == 0:
        json.dump(dict(statistics=stats, newton_its=drv.newton_its, linear_its=drv.lin_its), open(out_file, "w"))
    drv.ctx.close()


if __name__ == "__main__":
    main()
