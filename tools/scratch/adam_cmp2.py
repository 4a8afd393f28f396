import sys, os
sys.path.insert(0, os.getcwd())
import torch
sys.path.insert(0, "tests")
import test_train_step_gpu as T
a, b = T._scene(optimizer="torch", merged_render=True), T._scene(optimizer="fused", merged_render=True)
batches = T._batches(a, 3)
gen = torch.Generator().manual_seed(3)
msgs = [a.new_message(gen) for _ in range(3)]
for i, m in enumerate(msgs):
    ta0 = [e.weight.detach().clone() for e in a.model.msg_encoder.embeddings]
    tb0 = [e.weight.detach().clone() for e in b.model.msg_encoder.embeddings]
    la = a.train_step(batches[i % 3], m)
    lb = b.train_step(batches[i % 3], m)
    print(i, [f"{float(x):.8f}" for x in la], [f"{float(x):.8f}" for x in lb])
    bits = [int(v) for v in m.tolist()]
    sel = [2 * j + bit for j, bit in enumerate(bits)]
    ga = a.model.msg_encoder.embeddings[sel[0]].weight.grad
    Gb = b.optimizer.G
    print("   grad a (unscaled by scaler.unscale_?) sum/absmax/nnz:", float(ga.double().sum()), float(ga.abs().max()), int((ga != 0).sum()))
    print("   G b                              sum/absmax/nnz:", float(Gb.double().sum()), float(Gb.abs().max()), int((Gb != 0).sum()))
    for t in sel[:2]:
        da = a.model.msg_encoder.embeddings[t].weight.detach() - ta0[t]
        db = b.model.msg_encoder.embeddings[t].weight.detach() - tb0[t]
        diff = (da - db).abs()
        print(f"   table {t}: update a absmax {float(da.abs().max()):.3e} b {float(db.abs().max()):.3e}  nnz a {int((da!=0).sum())} b {int((db!=0).sum())}  max|da-db| {float(diff.max()):.3e}  n(|da-db|>1e-6) {int((diff>1e-6).sum())}")
    print("   scale a/b", float(a.scaler.get_scale()), float(b.scaler.get_scale()))
