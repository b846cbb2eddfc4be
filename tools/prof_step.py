"""Run a few un-graphed VAE train steps at BASELINE configs[1] (for ncu captures: `ncu ... python tools/prof_step.py 2`)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
scenes = int(sys.argv[2]) if len(sys.argv) > 2 else 64
syn = importlib.import_module("sln_b200.data.synthetic")
Model = importlib.import_module("sln_b200.models.Sg2ScVAE_model").Sg2ScVAEModel
sutils = importlib.import_module("sln_b200.utils")
dev = torch.device("cuda:0")
torch.manual_seed(42)
model = Model(syn.default_vocab(), embedding_dim=64, batch_size=128, train_3d=True, decoder_cat=True, gconv_mode='feedforward',
              gconv_num_layers=5, mlp_normalization='batch', vec_noise_dim=0, layout_noise_dim=32, use_AE=False).float().to(dev).train()
_, objs, boxes, triples, angles, attrs, _, _ = syn.synthetic_batch(scenes, 32, seed=42)
step = sutils.VAETrainStep(model, objs.size(0), triples.size(0), use_graph=False)
step.load_batch((objs, triples, boxes, angles, attrs))
for _ in range(n_steps):
    step.run()
torch.cuda.synchronize()
print("losses", step.losses.tolist())
