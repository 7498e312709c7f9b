"""Multi-GPU check of the forest path, launched by torchrun (one rank per GPU, NCCL):
pf_create_forest_distributed on the reference's adaptive goldens --
  * KAT-5, tests/hetero_3d_1.mpirun-4.statistics (BASELINE config 5 in small: 3-D, hanging nodes, per-cell E);
  * tests/miehe_shear_1.statistics (BASELINE config 4 in small: stress split + predictor-corrector AMR).
Prints MGPU_FOREST_OK on rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import cracks_b200 as pf
    from forest_cases import hetero_driver, miehe_forest_driver

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def fresh_id():
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(pf.PhaseFieldContext.nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        return idt.cpu().numpy().tobytes()

    d = (rank, world, fresh_id)
    close = lambda a, b, rel: abs(a - b) <= rel * abs(b)
    drv, g = hetero_driver(pf, d, device=local)
    stats = drv.run()
    if rank == 0:
        for got, ref in zip(stats, g["statistics"]):
            assert got["dofs"] == 5288 and close(got["crack"], ref["crack"], 1e-7) and close(got["bulk"], ref["bulk"], 1e-6), (got, ref)
        print("hetero_3d_1 on %d GPUs:" % world, [(s["bulk"], s["crack"]) for s in stats], flush=True)
    drv.ctx.close()

    g = json.load(open(os.path.join(ROOT, "tests", "golden", "miehe_shear_1.json")))
    drv = miehe_forest_driver(pf, g, dist=d, device=local)
    stats = drv.run()
    if rank == 0:
        assert [r["dofs"] for r in stats] == [r["dofs"] for r in g["statistics"]]
        for got, ref in zip(stats, g["statistics"]):
            tol = 1e-6 if got["step"] <= 9 else 1e-4
            for k in ("bulk", "crack", "load"):
                assert close(got[k], ref[k], tol), (got["step"], k, got[k], ref[k])
        print("miehe_shear_1 (adaptive) on %d GPUs: dofs" % world, [r["dofs"] for r in stats], flush=True)
        print("MGPU_FOREST_OK", flush=True)
    drv.ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
