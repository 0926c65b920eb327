"""``add_qspec_args``: the command-line surface of the plugin.

Flag names, defaults and value types follow the reference's ``training_args.py:36-256``
(they are what its example drivers and sweep scripts pass), declared here as one table.
One deliberate difference: ``--error`` stays a string (the reference parses it in argparse
and again in ``get_qconfig``, which raises at HEAD); both forms are accepted downstream.
"""
import argparse

__all__ = ["add_qspec_args", "SLURM_ARGS"]

QSPEC_HELP = (
    "Comma-separated quantization spec; the first field is the dtype, the rest are key=value "
    "with full names or abbreviations: qs=qscheme, qmax=quant_max, qmin=quant_min, "
    "ahl=amax_history_len, ax=ch_axis, bs=block_size. dtypes: intN, uintN, e4m3, e5m2, "
    "fp8_e4m3, fp8_e5m2, fp6_e3m2, fp6_e2m3, fp4_e2m1, positN_ES. "
    "Example: int8,qs=per_tensor_symmetric,qmax=127,ahl=50"
)

SLURM_ARGS = {
    "job-name": {"type": str, "default": "test"},
    "partition": {"type": str, "default": "gpu"},
    "nodes": {"type": int, "default": 1},
    "time": {"type": str, "default": "48:00:00"},
    "gpus": {"type": str, "default": "1"},
    "cpus": {"type": int, "default": 8},
    "mem": {"type": str, "default": "16GB"},
    "output": {"type": str, "default": None},
    "error": {"type": str, "default": None},
    "exclude": {"type": str, "default": None},
    "nodelist": {"type": str, "default": None},
}


def _csv(text):
    return text.split(",")


_FLAG = "store_true"
# (flag, kwargs) in the reference's order: logging, training, quantization
_OPTIONS = [
    ("--project", dict(default=None, help="W&B project the run is sent to.")),
    ("--run_name", dict(default=None, help="Display name of the run.")),
    ("--run_id", dict(default=None, help="W&B run id, used for resuming.")),
    ("--sweep_config", dict(default=None, help="JSON file with a W&B sweep configuration.")),
    ("--sweep_id", dict(default=None, help="Identifier of an existing W&B sweep.")),
    ("--max_trials", dict(type=int, default=None, help="Number of sweep trials to run.")),
    ("--log_level", dict(choices=["DEBUG", "INFO", "WARNING", "ERROR", "CRITICAL"], default="WARNING",
                         help="Logging level.")),
    ("--log_file", dict(default=None, help="Log file; stdout when omitted.")),
    ("--gpu", dict(type=int, default=None, help="GPU to use.")),
    ("--do_train", dict(action=_FLAG, help="Run training.")),
    ("--sgd", dict(action=_FLAG, help="Use the SGD optimizer.")),
    ("--warmup_ratio", dict(type=float, default=0.0, help="Warm-up fraction of the lr schedule.")),
    ("--bf16", dict(action=_FLAG, help="Run the model in bfloat16.")),
    ("--num_hidden_layers", dict(type=int, default=None, help="Number of encoder layers to keep.")),
    ("--lora_rank", dict(type=int, default=0, help="Rank of the LoRA update matrices.")),
    ("--lora_alpha", dict(type=int, default=8, help="LoRA scaling factor.")),
    ("--target_modules", dict(type=_csv, default="query,value", help="Modules that receive LoRA updates.")),
    ("--peft_model_id", dict(default=None, help="Pre-trained PEFT adapter.")),
    ("--pt2e", dict(action=_FLAG, help="Use the torch.export post-training quantization flow.")),
    ("--activation", dict(default=None, help="Activation quantization spec. " + QSPEC_HELP)),
    ("--output_activation", dict(default=None, help="Output-activation quantization spec.")),
    ("--weight", dict(default=None, help="Weight quantization spec.")),
    ("--bias", dict(default=None, help="Bias quantization spec.")),
    ("--error", dict(default=None, help="Activation-gradient quantization spec.")),
    ("--quantize_forward", dict(default="gemm", help="Forward op groups to quantize: gemm, residual, "
                                                     "activation, layernorm, scaling (comma separated).")),
    ("--quantize_backprop", dict(default="gemm", help="Backward op groups to quantize (same choices).")),
    ("--force_scale_power_of_two", dict(action=_FLAG, help="Round scaling factors up to a power of two.")),
    ("--calibration_steps", dict(type=int, default=0, help="Calibration steps for PTQ.")),
    ("--convert_model", dict(action=_FLAG, help="Convert the model to a quantized model.")),
    ("--compile", dict(action=_FLAG, help="Generate an accelerator program for the model.")),
    ("--op_fusion", dict(type=_csv, default=None, help="Module-name substrings whose inputs stay unquantized "
                                                       "(the op is fused with the previous GEMM).")),
    ("--posit_exp", dict(action=_FLAG, help="Posit-approximated exp in softmax.")),
    ("--posit_exp_shifted", dict(action=_FLAG, help="Shifted posit-approximated exp in softmax.")),
    ("--posit_reciprocal", dict(action=_FLAG, help="Posit-approximated reciprocal in softmax.")),
    ("--record_histogram", dict(action=_FLAG, help="Record exponent histograms of quantized tensors.")),
    ("--bank_width", dict(type=int, default=None, help="Memory bank width in bytes (accelerator flow).")),
]


def add_qspec_args(parser=None):
    if parser is None:
        parser = argparse.ArgumentParser(description="Run quantized inference or training.")
    for flag, kwargs in _OPTIONS:
        parser.add_argument(flag, **kwargs)
    sub = parser.add_subparsers(help="sub-command help", dest="action")
    slurm = sub.add_parser("slurm", help="slurm command help")
    for name, kwargs in SLURM_ARGS.items():
        slurm.add_argument("--" + name, **kwargs)
    sub.add_parser("bash", help="bash command help")
    return parser
