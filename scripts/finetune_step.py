"""Quantized fine-tune step (BASELINE configs[3]): RoBERTa-base shape, LoRA r=8 on query,value + classifier head,
E4M3 forward (activations + weights) / E5M2 per-tensor delayed-scaling gradients, op groups gemm,residual,layernorm,
activation in both directions (run_quantized_training.py:213-235 of the reference; the loop is
examples/text_classification/run_glue_no_trainer.py:658-668), AdamW, grad-clip 1.0.  Data parallel: one process per
GPU; the TRAINABLE gradients are all-reduced (NCCL, averaged) by quantized_training.dp.GradReducer from grad hooks,
i.e. overlapped with the rest of the backward pass.  Forward GEMMs, dgrad and wgrad all run on the tcgen05 kernel.

    python scripts/finetune_step.py [--steps 20] [--graph]            # 1 GPU
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/finetune_step.py

--graph captures the whole step (forward, backward, all-reduce, clip, AdamW) in ONE CUDA graph: the eager step is
host-bound (~600 fake-quant module calls and hooks per step).  Random-init weights, synthetic batch [16, 128] per GPU
(no network).  Prints one JSON line (rank 0).  bench.py imports `run()` for its `finetune` record."""
import argparse, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantized-training_b200"))
import quantized_training as qt
from quantized_training import dp
from quantized_training.modules.lora import apply_lora

ERROR = "fp8_e5m2,qs=per_tensor_symmetric,qmax=57344,ahl=10"
OPS = "gemm,residual,layernorm,activation"


def build(dev, rank, layers=12, batch=16, seq=128, activation="fp8_e4m3", error=ERROR):
    from transformers import RobertaConfig, RobertaForSequenceClassification
    torch.manual_seed(0)                                   # identical initial weights on every rank
    cfg = RobertaConfig(vocab_size=50265, max_position_embeddings=514, type_vocab_size=1, num_labels=3,
                        num_hidden_layers=layers, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = RobertaForSequenceClassification(cfg).to(dev)
    apply_lora(model, ["query", "value"], r=8, lora_alpha=8)
    for p in model.classifier.parameters():                # peft modules_to_save=["classifier"]
        p.requires_grad_(True)
    args = qt.add_qspec_args().parse_args([
        "--activation", activation, "--weight", activation, "--error", error,
        "--quantize_forward", OPS, "--quantize_backprop", OPS, "--bf16", "--do_train"])
    qt.quantize(model, args)
    model.train()
    torch.manual_seed(1234 + rank)                         # each rank its own shard of the batch
    ids = torch.randint(3, cfg.vocab_size, (batch, seq), device=dev)
    labels = torch.randint(0, 3, (batch,), device=dev)
    # warm-up forward/backward: the hook fake-quantizers are created lazily, before the optimizer (reference
    # run_glue_no_trainer.py:471-474)
    model(input_ids=ids, labels=labels).loss.backward()
    model.zero_grad(set_to_none=True)
    return model, ids, labels


def run(dev, rank, world, dist, steps=20, warmup=3, graph=False, allreduce=True, **kw):
    """Times `steps` optimizer steps; returns a dict (ms_per_step is the MAX over ranks)."""
    model, ids, labels = build(dev, rank, **kw)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1.4e-3, weight_decay=0.0, capturable=graph, foreach=True)
    red = dp.GradReducer(params) if allreduce else None
    if red is None:
        model.zero_grad(set_to_none=False)
    loss_buf = torch.zeros((), device=dev)

    def step():
        if red is not None:
            red.zero_grad()
        else:
            for p in params:
                if p.grad is not None:
                    p.grad.zero_()
        loss = model(input_ids=ids, labels=labels).loss
        loss.backward()
        nb = red.finish() if red is not None else 0
        torch.nn.utils.clip_grad_norm_(params, 1.0, foreach=True)   # no host read-back (error_if_nonfinite off)
        opt.step()
        loss_buf.copy_(loss.detach())
        return nb

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    losses = []
    nb = 0
    for _ in range(max(warmup, 3)):
        nb = step()
        losses.append(float(loss_buf))
    replay = step
    if graph:
        g = torch.cuda.CUDAGraph()
        barrier()
        with torch.cuda.graph(g):
            step()
        replay = g.replay
        replay()
        losses.append(float(loss_buf))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        replay()
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    losses.append(float(loss_buf))
    ms = dp.reduce_max(e0.elapsed_time(e1) / steps)
    nfq = sum(1 for m in model.modules() if isinstance(m, qt.FusedAmaxObsFakeQuantize))
    err_fq = [m for n, m in model.named_modules() if "error_pre_process" in n and isinstance(m, qt.FusedAmaxObsFakeQuantize)]
    batch, seq = ids.shape
    out = {
        "workload": f"RoBERTa-base shape ({kw.get('layers', 12)} layers) LoRA r=8 fine-tune step, batch {batch} x seq {seq} "
                    f"per GPU, {kw.get('activation', 'fp8_e4m3')} forward / {kw.get('error', ERROR)} gradients, ops {OPS}; "
                    "forward, dgrad and wgrad GEMMs on the tcgen05 kernel",
        "mode": "whole step in one CUDA graph" if graph else "eager",
        "n_gpus": world, "steps": steps, "ms_per_step": ms, "sequences_per_s": world * batch / ms * 1e3,
        "wall_ms_per_step": wall / steps * 1e3,
        "trainable_params": sum(p.numel() for p in params),
        "allreduce": ("NCCL all-reduce(AVG) of the trainable gradients, launched from grad hooks during backward "
                      f"({nb} bucket(s))") if (allreduce and world > 1) else ("none (1 GPU)" if allreduce else "disabled"),
        "fake_quant_modules": nfq, "gradient_fake_quant_modules": len(err_fq),
        "gradient_scale_example": float(err_fq[0].scale) if err_fq else None,
        "loss_first": losses[0], "loss_last": losses[-1]}
    if red is not None:
        red.remove()
    del model, opt
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--seq", type=int, default=128)
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--activation", default="fp8_e4m3")
    ap.add_argument("--error", default=ERROR)
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--no-allreduce", action="store_true")
    a = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    out = run(dev, rank, world, dist, steps=a.steps, warmup=a.warmup, graph=a.graph, allreduce=not a.no_allreduce,
              layers=a.layers, batch=a.batch, seq=a.seq, activation=a.activation, error=a.error)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
