"""Eager front end: ``quantize(model, args)`` (reference: quantize.py:52-283).

One-time model surgery; afterwards every forward/backward runs the fake-quant kernels from hooks:
  1. HF attention/output blocks -> quantizable blocks (hookable matmul / scaling / softmax / residual)
  2. optional bf16 cast
  3. QConfig from the three qspec strings, propagated to every module
  4. nn.Linear / LoRA linear -> QAT modules (weight fake-quantized each forward)
  5. hooks: forward-pre (activations), full-backward-pre (gradients w.r.t. outputs),
     full-backward (gradients w.r.t. inputs of residual-feeding layers)
Which modules get hooks is chosen by op-group names: gemm, residual, layernorm, activation, scaling.
"""
import copy
import logging

import torch
import torch.ao.nn.intrinsic as nni
import torch.nn as nn
from torch.nn.utils.parametrize import type_before_parametrizations
from transformers import PretrainedConfig

from .modules.quantizable.functional_modules import MatmulFunctional
from .qconfig import get_qconfig
from .quantization_mappings import (
    DEFAULT_QAT_MODULE_MAPPINGS,
    QCONFIG_PROPAGATE_MODULE_CLASS_LIST,
    TRANSFORMER_MODULE_MAPPINGS,
)

__all__ = ["propagate_config", "quantize", "prepare", "convert", "replace_softmax", "get_quantized_model"]

logger = logging.getLogger(__name__)

# Linear layers whose *input* gradient feeds a residual sum in the backward pass
RESIDUAL_LAYERS_BWD = [
    "attention.self.query",
    "attention.self.key",
    "attention.self.value",
    "intermediate.dense",
    "bottleneck.input.dense",
    "bottleneck.attention.dense",
]
_HOOK_KINDS = ("activation_pre_process", "error_pre_process", "error_post_process")


def propagate_config(module, name, qconfig):
    """setattr(name, qconfig) on `module` and every descendant."""
    for m in module.modules():
        setattr(m, name, qconfig)


def quantize(model, args, inplace=True):
    if not inplace:
        model = copy.deepcopy(model)

    if getattr(args, "posit_exp", False) or getattr(args, "posit_exp_shifted", False) \
            or getattr(args, "posit_reciprocal", False):
        replace_softmax(model, args.posit_exp, args.posit_exp_shifted, args.posit_reciprocal)

    wants_hooks = args.activation is not None or args.error is not None
    if wants_hooks and isinstance(getattr(model, "config", None), PretrainedConfig):
        propagate_config(model, "config", model.config)
        convert(model, inplace=True, custom_module_class_mapping=TRANSFORMER_MODULE_MAPPINGS)

    if hasattr(model, "hf_device_map"):
        try:
            from accelerate import dispatch_model
        except ImportError as e:  # layer placement across GPUs is outside the data-parallel hot path
            raise RuntimeError("model.hf_device_map is set but accelerate is not installed") from e
        dispatch_model(model, device_map=model.hf_device_map)

    if getattr(args, "bf16", False):
        model.bfloat16()

    # the reference's drivers read these back
    if args.activation is None:
        args.quantize_forward = None
    if args.error is None:
        args.quantize_backprop = None

    qconfig = get_qconfig(args.activation, args.weight, args.error,
                          getattr(args, "record_histogram", False),
                          getattr(args, "force_scale_power_of_two", False))
    propagate_config(model, "qconfig", qconfig)
    convert(model, mapping=DEFAULT_QAT_MODULE_MAPPINGS, inplace=True)
    prepare(model, True, args.quantize_forward, args.quantize_backprop, getattr(args, "op_fusion", None))
    return model


def _parse_ops(op_str):
    groups = QCONFIG_PROPAGATE_MODULE_CLASS_LIST
    ops = {op.lower() for op in op_str.split(",")} if op_str is not None else set()
    invalid_ops = ops - set(groups)
    assert not invalid_ops, f"Invalid operation(s) {', '.join(invalid_ops)}. Options are {', '.join(groups)}."
    return tuple(cls for op in ops for cls in groups[op])


def _get_unique_devices_(mod):
    return {p.device for p in mod.parameters()} | {b.device for b in mod.buffers()}


def _register_module_hook(module, hook_name, name):
    """Attach a ModuleDict of lazily created fake-quantizers (one per positional tensor argument)
    and the hook that applies them."""
    assert hook_name in _HOOK_KINDS
    fq_by_arg = nn.ModuleDict()
    module.add_module(hook_name, fq_by_arg)
    make_fq = module.qconfig.activation if hook_name == "activation_pre_process" else module.qconfig.error

    def quantize_args(mod, tensors):
        out = []
        for i, t in enumerate(tensors):
            if isinstance(t, torch.Tensor):
                key = str(i)
                if key not in fq_by_arg:  # created on first use, on the tensor's device
                    fq = make_fq(device=t.device)
                    fq.name = f"{name}.{key}"
                    if hook_name == "activation_pre_process" and isinstance(module, MatmulFunctional) \
                            and hasattr(fq, "preserve_strides"):
                        fq.preserve_strides = True  # k^T stays a view: ops.matmul reads it K-major, no copy
                    fq_by_arg[key] = fq
                t = fq_by_arg[key](t)
            out.append(t)
        return tuple(out)

    if hook_name == "activation_pre_process":
        module.register_forward_pre_hook(quantize_args)
    elif hook_name == "error_pre_process":
        module.register_full_backward_pre_hook(quantize_args)
    else:
        module.register_full_backward_hook(lambda mod, grad_inputs, grad_outputs: quantize_args(mod, grad_inputs))


def _add_observer_(module, fwd_classes, bwd_classes, bwd_residual, op_fusion, prefix):
    residual_classes = _parse_ops("residual")

    def visit(m, name):
        if getattr(m, "qconfig", None) is None:
            return
        if op_fusion is not None and any(tag in name for tag in op_fusion):
            return  # fused with the producing GEMM: its inputs stay in high precision
        if isinstance(m, fwd_classes):
            _register_module_hook(m, "activation_pre_process", name)
        if isinstance(m, bwd_classes):
            _register_module_hook(m, "error_pre_process", name)
        if bwd_residual and (any(tag in name for tag in RESIDUAL_LAYERS_BWD) or isinstance(m, residual_classes)):
            _register_module_hook(m, "error_post_process", name)

    for child_name, child in list(module.named_children()):
        child_prefix = f"{prefix}.{child_name}" if prefix else child_name
        if isinstance(child, nni._FusedModule):
            visit(child, child_prefix)
        else:
            _add_observer_(child, fwd_classes, bwd_classes, bwd_residual, op_fusion, child_prefix)
    visit(module, prefix)


def prepare(model, inplace=False, fwd_quantized_ops=None, bwd_quantized_ops=None, op_fusion=None):
    if not inplace:
        model = copy.deepcopy(model)
    bwd_residual = bool(bwd_quantized_ops) and "residual" in bwd_quantized_ops
    _add_observer_(model, _parse_ops(fwd_quantized_ops), _parse_ops(bwd_quantized_ops), bwd_residual,
                   op_fusion, prefix="")
    return model


def convert(module, mapping=None, inplace=False, custom_module_class_mapping=None):
    """Swap child modules by type: `custom_module_class_mapping` via ``from_observed`` (whole blocks),
    `mapping` via ``from_float`` (modules carrying a qconfig)."""
    if not inplace:
        module = copy.deepcopy(module)
    _convert(module, DEFAULT_QAT_MODULE_MAPPINGS if mapping is None else mapping,
             custom_module_class_mapping or {})
    return module


def _convert(module, mapping, custom):
    swapped = {}
    for name, child in module.named_children():
        whole_unit = isinstance(child, nni._FusedModule) or type_before_parametrizations(child) in custom
        if not whole_unit:
            _convert(child, mapping, custom)
        swapped[name] = swap_module(child, mapping, custom)
    for name, new_child in swapped.items():
        module._modules[name] = new_child
    return module


def swap_module(mod, mapping, custom_module_class_mapping):
    kind = type_before_parametrizations(mod)
    if kind in custom_module_class_mapping:
        new_mod = custom_module_class_mapping[kind].from_observed(mod)
        # blocks nested inside a swapped block (e.g. attention inside a decoder layer) keep converting
        _convert(new_mod, mapping, custom_module_class_mapping)
    elif getattr(mod, "qconfig", None) is not None and kind in mapping:
        new_mod = mapping[kind].from_float(mod)
    else:
        return mod
    # hooks registered on the float module keep firing on its replacement
    for fn in mod._forward_pre_hooks.values():
        new_mod.register_forward_pre_hook(fn)
    for fn in mod._forward_hooks.values():
        new_mod.register_forward_hook(fn)
    for fn in mod._backward_pre_hooks.values():
        new_mod.register_full_backward_pre_hook(fn)
    for fn in mod._backward_hooks.values():
        new_mod.register_full_backward_hook(fn)
    devices = _get_unique_devices_(mod)
    assert len(devices) <= 1, f"swap_module only works with cpu or single-device CUDA modules, but got devices {devices}"
    if devices:
        new_mod.to(next(iter(devices)))
    return new_mod


def replace_softmax(module, posit_exp, posit_exp_shifted, posit_reciprocal, dtype=None, device=None):
    raise NotImplementedError(
        "posit-approximated softmax needs the reference's posit16 gold tables (missing blobs) and is outside "
        "the B200 hot path (SURVEY.md §2 row 14)")


def get_quantized_model(model, qconfig, op_fusion=None, device=None):
    raise NotImplementedError(
        "get_quantized_model is the reference's legacy full-model-copy path, superseded by quantize(); "
        "outside the B200 hot path (SURVEY.md §2 row 15)")
