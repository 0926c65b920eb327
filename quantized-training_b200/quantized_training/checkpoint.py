"""Training-state checkpoints in the layout the reference's drivers write (reference:
examples/question_answering/run_qa_no_trainer.py:961-990 `save_state` / `load_state`, :1021-1056 resume parsing).

One ``checkpoint.tar`` per ``step_N`` / ``epoch_N`` folder holding ``model_state_dict`` (quantizer buffers included:
scale, amax_history, the exponent histogram), ``optimizer_state_dict``, ``scheduler_state_dict``, ``best_metric`` and
``run_id``.  The quantizers' lazily shaped buffers are resized on load (fake_quantize.py `_load_from_state_dict`), so a
freshly prepared model can take a state saved after calibration.  The evaluation / wandb / `save_pretrained` parts of
the reference's `save_state` stay with the driver.
"""
import os
import re
from typing import Optional, Tuple

import torch

__all__ = ["save_state", "load_state", "find_latest", "parse_resume"]

FILE_NAME = "checkpoint.tar"


def save_state(output_dir, model, optimizer=None, lr_scheduler=None, best_metric=None, run_id=None):
    """Write ``output_dir/checkpoint.tar``; returns the path."""
    os.makedirs(output_dir, exist_ok=True)
    path = os.path.join(output_dir, FILE_NAME)
    torch.save({
        "model_state_dict": model.state_dict(),
        "optimizer_state_dict": optimizer.state_dict() if optimizer is not None else None,
        "scheduler_state_dict": lr_scheduler.state_dict() if lr_scheduler is not None else None,
        "best_metric": best_metric,
        "run_id": run_id,
    }, path)
    return path


def load_state(output_dir, model, optimizer=None, lr_scheduler=None, map_location=None):
    """Restore what `save_state` wrote and return the whole checkpoint dict (``best_metric``, ``run_id``...)."""
    checkpoint = torch.load(os.path.join(output_dir, FILE_NAME), map_location=map_location, weights_only=False)
    model.load_state_dict(checkpoint["model_state_dict"])
    if optimizer is not None and checkpoint.get("optimizer_state_dict") is not None:
        optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
    if lr_scheduler is not None and checkpoint.get("scheduler_state_dict") is not None:
        lr_scheduler.load_state_dict(checkpoint["scheduler_state_dict"])
    return checkpoint


def find_latest(root) -> Optional[str]:
    """Most recently modified ``step_N`` / ``epoch_N`` folder under `root` that holds a checkpoint."""
    dirs = [os.path.join(root, d) for d in os.listdir(root)
            if re.fullmatch(r"(step|epoch)_\d+", d) and os.path.isfile(os.path.join(root, d, FILE_NAME))]
    return max(dirs, key=os.path.getmtime) if dirs else None


def parse_resume(path, steps_per_epoch, gradient_accumulation_steps=1) -> Tuple[int, Optional[int], int]:
    """``(starting_epoch, resume_step, completed_steps)`` from a ``step_N`` / ``epoch_N`` folder name, the way the
    reference's loop restarts (`steps_per_epoch` = len(train_dataloader); `resume_step` counts batches to skip inside
    the starting epoch, None when restarting on an epoch boundary)."""
    name = os.path.splitext(os.path.basename(os.path.normpath(path)))[0]
    m = re.fullmatch(r"(step|epoch)_(\d+)", name)
    if m is None:
        raise ValueError(f"not a step_N / epoch_N checkpoint folder: {path!r}")
    n = int(m.group(2))
    updates_per_epoch = -(-steps_per_epoch // gradient_accumulation_steps)
    if m.group(1) == "epoch":
        return n + 1, None, (n + 1) * updates_per_epoch
    resume_step = n * gradient_accumulation_steps
    starting_epoch = resume_step // steps_per_epoch
    return starting_epoch, resume_step - starting_epoch * steps_per_epoch, n
