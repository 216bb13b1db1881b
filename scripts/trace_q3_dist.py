"""per-launch times of Q3 on N GPUs (sharded): rank 0 prints the engine trace"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
from resql_b200 import Engine, Plan, native as N
from resql_b200 import tpch_device as TD
from common import load_plan_dict
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("gloo")
eng = Engine(rank)
uid = [eng.dist_unique_id() if rank == 0 else None]; dist.broadcast_object_list(uid, src=0); eng.dist_init(rank, world, uid[0])
orders, li, cust = TD.gen_orders_lineitem(100, 42, dev, rank=rank, world=world)
src = {"lineitem": li, "orders": orders, "customer": cust}
d = load_plan_dict("q3")
tabs = {t["name"]: eng.upload_device(t["name"], TD.as_device_columns(src[t["name"]], list(src[t["name"]].keys())), src[t["name"]][t["columns"][0]].shape[0], borrow=False) for t in d["tables"]}
plan = Plan(d)
for i in range(5):
    if i == 4:
        eng.set_option("graphs", 0); eng.set_option("replay", 0); eng.set_option("trace", 1)
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    res, tm = eng.execute(plan, tabs, N.RQ_PLAN_SHARDED)
    if rank == 0:
        print("q3 wall_ms", round(1e3 * (time.perf_counter() - t0), 3), "kernel_ms", round(tm.kernel_ms, 3), "nccl_ms", round(tm.nccl_ms, 3), "syncs", tm.host_syncs, flush=True)
dist.barrier()
