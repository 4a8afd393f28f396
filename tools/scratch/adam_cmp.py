import sys, os
sys.path.insert(0, os.getcwd())
import torch
sys.path.insert(0, "tests")
import test_train_step_gpu as T
a, b = T._scene(optimizer="torch"), T._scene(optimizer="fused")
batches = T._batches(a, 3)
gen = torch.Generator().manual_seed(3)
msgs = [a.new_message(gen) for _ in range(6)]
for i, m in enumerate(msgs):
    la = a.train_step(batches[i % 3], m)
    lb = b.train_step(batches[i % 3], m)
    print(i, [f"{float(x):.8f}" for x in la], [f"{float(x):.8f}" for x in lb], f"rel {abs(float(la[0])-float(lb[0]))/abs(float(la[0])):.2e}")
