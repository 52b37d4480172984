import sys, os
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np, torch
from test_finetune_gpu import finetune_case, feed, rng_with_masks
from helpers import O, engine_for
B, T = 2, 5
cfg, params, batch, keep = finetune_case(B, T, seed=B)
eng = engine_for(cfg, params, B, T)
dev = lambda a, dt: torch.tensor(np.ascontiguousarray(a)).to(dt).cuda()
f = feed(batch)
eng.forward_backward_device(dev(f["image_f_inputs"], torch.float32), dev(f["ann_inputs_enc"], torch.int32),
                            dev(f["ann_inputs_dec"], torch.int32), dev(f["ann_lengths"], torch.int32), 0,
                            rng=rng_with_masks(batch, keep))
torch.cuda.synchronize()
res, grads, gnorm = O.compute_grads(params, cfg, batch, emulate=True)
out = {}
for n in ["cnn/fc2/biases", "cnn/fc1/biases", "cnn/conv5_3/biases_conv", "cnn/conv1_1/weights", "cnn/conv1_1/biases",
          "cnn/conv5_3/weights_conv", "cnn/conv1_2/biases", "imf_emb/bias"]:
    out["got:" + n] = eng.get_gradient(n)
    out["ref:" + n] = grads[n].numpy()
np.savez("gpurun_out/ft_dump.npz", **out)
