"""torchrun --nproc-per-node N tools/train_ddp.py [steps] [graph|eager]: data-parallel training on the synthetic scene;
checks that the replicas stay bit-identical (same averaged gradients -> same Adam updates) and reports throughput."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist
from fruitnerf_b200.scripts.train import synthetic_spec
from fruitnerf_b200.trainer import Trainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
graph = (sys.argv[2] if len(sys.argv) > 2 else "graph") == "graph"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(0)  # identical initial parameters on every rank (DDP broadcasts rank 0's; same seed is equivalent here)
tr = Trainer(synthetic_spec("fruit_nerf", schedule_steps=steps), device=f"cuda:{local}", world_size=world, local_rank=rank, use_cuda_graph=graph)
tr.train(20)
torch.cuda.synchronize(); dist.barrier()
t0 = time.time()
hist = tr.train(steps - 20, log_every=steps - 20, eval_every=10**9)
torch.cuda.synchronize(); dist.barrier()
dt = time.time() - t0
table = tr.pipeline.model.field.mlp_base_grid.hash_table
chk = torch.stack([table.double().sum(), tr.pipeline.model.field.mlp_head.layers[0].weight.double().sum(),
                   tr.pipeline.model.proposal_networks[0].encoding.hash_table.double().sum()])
gathered = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(gathered, chk)
same = all(torch.equal(g, gathered[0]) for g in gathered)
if rank == 0:
    rays = tr.spec.pipeline.datamanager.train_num_rays_per_batch
    print({"world": world, "graph": graph, "ms_per_iteration": 1e3 * dt / (steps - 20), "rays_per_s_total": world * rays * (steps - 20) / dt,
           "replicas_identical": same, "loss": hist[-1]["loss"], "psnr": hist[-1]["psnr"]})
dist.destroy_process_group()
