"""Print the quantities tests/test_gpu_training.py asserts on (to see the margins)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from tests.test_gpu_training import _tiny_spec
from fruitnerf_b200.trainer import Trainer
torch.manual_seed(0)
tr = Trainer(_tiny_spec(), device="cuda:0", use_cuda_graph=True)
h = tr.train(400, log_every=50, eval_every=10**9)
print("loss0", h[0]["loss"], "lossN", h[-1]["loss"], "psnr0", h[0]["psnr"], "psnrN", h[-1]["psnr"])
print("eval", tr.pipeline.get_average_eval_image_metrics(tr.step), "prop steps", tr.optimizers["proposal_networks"].step_count)
torch.manual_seed(0)
e = Trainer(_tiny_spec(), device="cuda:0", use_cuda_graph=False)
he = e.train(100, log_every=50, eval_every=10**9)
print("eager", he[0]["loss"], he[-1]["loss"], he[-1]["psnr"])
