"""Quantized fine-tune step (BASELINE configs[3]): RoBERTa-base shape, LoRA r=8 on query,value + classifier head,
E4M3 forward (activations + weights) / E5M2 per-tensor delayed-scaling gradients, op groups gemm,residual,layernorm,
activation in both directions (run_quantized_training.py:213-235 of the reference), AdamW, grad-clip 1.0; data parallel:
one process per GPU, all-reduce(SUM)/world of the TRAINABLE gradients only (quantized_training.dp).

    python scripts/finetune_step.py [--steps 20]                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/finetune_step.py

Random-init weights, synthetic batch [16, 128] per GPU (no network).  Prints one JSON line (rank 0)."""
import argparse, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantized-training_b200"))
import quantized_training as qt
from quantized_training import dp
from quantized_training.modules.lora import apply_lora


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--seq", type=int, default=128)
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--activation", default="fp8_e4m3")
    ap.add_argument("--error", default="fp8_e5m2,qs=per_tensor_symmetric,qmax=57344,ahl=10")
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
    from transformers import RobertaConfig, RobertaForSequenceClassification
    torch.manual_seed(0)                                   # identical initial weights on every rank
    cfg = RobertaConfig(vocab_size=50265, max_position_embeddings=514, type_vocab_size=1, num_labels=3,
                        num_hidden_layers=a.layers, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model = RobertaForSequenceClassification(cfg).to(dev)
    apply_lora(model, ["query", "value"], r=8, lora_alpha=8)
    for p in model.classifier.parameters():                # peft modules_to_save=["classifier"]
        p.requires_grad_(True)
    args = qt.add_qspec_args().parse_args([
        "--activation", a.activation, "--weight", a.activation, "--error", a.error,
        "--quantize_forward", "gemm,residual,layernorm,activation", "--quantize_backprop", "gemm,residual,layernorm,activation",
        "--bf16", "--do_train"])
    qt.quantize(model, args)
    model.train()
    torch.manual_seed(1234 + rank)                         # each rank its own shard of the batch
    ids = torch.randint(3, cfg.vocab_size, (a.batch, a.seq), device=dev)
    labels = torch.randint(0, 3, (a.batch,), device=dev)
    # warm-up forward/backward: the hook fake-quantizers are created lazily, before the optimizer (reference
    # run_glue_no_trainer.py:471-474)
    model(input_ids=ids, labels=labels).loss.backward()
    model.zero_grad(set_to_none=True)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1.4e-3, weight_decay=0.0)

    def step():
        loss = model(input_ids=ids, labels=labels).loss
        loss.backward()
        nb = dp.allreduce_grads_(params)
        torch.nn.utils.clip_grad_norm_(params, 1.0, error_if_nonfinite=True)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss.detach(), nb

    losses = []
    for _ in range(a.warmup):
        l, nb = step(); losses.append(float(l))
    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        l, nb = step(); losses.append(l)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / a.steps
    ms = dp.reduce_max(ms)
    losses = [float(x) for x in losses]
    nfq = sum(1 for m in model.modules() if isinstance(m, qt.FusedAmaxObsFakeQuantize))
    err_fq = [m for n, m in model.named_modules() if "error_pre_process" in n and isinstance(m, qt.FusedAmaxObsFakeQuantize)]
    if rank == 0:
        print(json.dumps({
            "workload": f"RoBERTa-base shape ({a.layers} layers) LoRA r=8 fine-tune step, batch {a.batch} x seq {a.seq} per GPU, "
                        f"{a.activation} forward / {a.error} gradients, ops gemm,residual,layernorm,activation",
            "n_gpus": world, "ms_per_step": ms, "sequences_per_s": world * a.batch / ms * 1e3,
            "trainable_params": sum(p.numel() for p in params), "allreduce_buckets": nb,
            "fake_quant_modules": nfq, "gradient_fake_quant_modules": len(err_fq),
            "gradient_scale_example": float(err_fq[0].scale) if err_fq else None,
            "loss_first": losses[0], "loss_last": losses[-1], "losses": [round(x, 4) for x in losses[:: max(1, len(losses) // 8)]]}), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
