"""Scratch (torchrun, N ranks): cost of qlc_stats_allreduce after every bench step, for QLC_COMM_RESERVE_SMS values (read once per process:
one torchrun per value)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
q = importlib.import_module("q-learning_b200")
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
env = q.BreakoutEnvironment(n_envs=4096, seed=1, env_id_base=rank * 4096, replay_capacity=1 << 20, device=local)
acts = torch.randint(0, 3, (64, 4096), dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream().cuda_stream
launch = lambda: env.step_device(acts.data_ptr(), 64, None, None, s)
def barrier():
    dist.barrier(); torch.cuda.synchronize()
for _ in range(50): launch()
r = bench.measure_stats_reduce_every_step(q, torch, dist, env, launch, s, rank, world, dev, barrier, 300)
if rank == 0:
    print("reserve=%s world=%d: with %.4f ms, without %.4f ms, overhead %.2f %%" % (os.environ.get("QLC_COMM_RESERVE_SMS", "default"), world, r["ms_per_step_with"], r["ms_per_step_without"], 100 * r["overhead_frac"]), flush=True)
env.close(); dist.barrier(); dist.destroy_process_group()
