"""On-disk compatibility with the reference's runs:

* ``.pth`` checkpoints: ``{"model_state_dict", "optimizer_state_dict", "step_count", "best_metrics",
  "scheduler_state_dict"}`` (trainer.py:599-612 ``Trainer.save_model``; resume :72-86 ``load_model_from_checkpoint``;
  utils/nbody_utils.py:1379-1413 ``load_checkpoint``; :1316-1376 ``load_model_for_inference``).  The model keys are the
  reference's parameter names; e3nn-internal entries of a reference checkpoint (``tp.output_mask``, compiled-module
  constants, the empty ``gate.mul.weight``) are dropped on load and ``tp.output_mask`` is re-created on save.
* run directory: ``model_params.json`` (trainer.py:526-528), ``<dataset_name>_dataset/metadata.json`` (:530-537;
  keys from dataset_gravity_otf.py:257-275 -- note ``n_balls``, which the dataset constructor calls ``num_nodes``),
  found again through ``get_dataset_metadata_path`` (nbody_utils.py:1446-1488) and
  ``load_dataset_from_metadata_file`` (visualization_utils.py:1438-1452).

Pure host-side file handling: nothing here touches the kernels.
"""
from __future__ import annotations

import inspect
import json
import os
import re
from typing import Optional

import torch

__all__ = ["save_model", "load_checkpoint", "load_model_for_inference", "save_run_metadata",
           "get_dataset_metadata_path", "load_dataset_from_metadata_file", "reference_state_dict"]


def reference_state_dict(model) -> dict:
    """``model.state_dict()`` plus the ``tp.output_mask`` buffer e3nn's TensorProduct keeps in its state (all ones:
    every output irrep of these products is reachable), so that the reference's strict ``load_state_dict`` finds it."""
    sd = dict(model.state_dict())
    for name, mod in model.named_modules():
        tp = getattr(mod, "tp", None)
        if tp is not None and hasattr(tp, "irreps_out") and hasattr(tp, "weight"):
            key = f"{name}.tp.output_mask" if name else "tp.output_mask"
            sd[key] = torch.ones(tp.irreps_out.dim, dtype=tp.weight.dtype, device=tp.weight.device)
    return sd


def save_model(model, optimizer=None, lr_scheduler=None, step_count: int = 0, best_metrics: Optional[dict] = None,
               save_path: str = ".", filename: str = "model.pth") -> str:
    """trainer.py:599-612."""
    checkpoint = {"model_state_dict": reference_state_dict(model),
                  "optimizer_state_dict": optimizer.state_dict() if optimizer is not None else {},
                  "step_count": step_count, "best_metrics": best_metrics if best_metrics is not None else {}}
    if lr_scheduler is not None:
        checkpoint["scheduler_state_dict"] = lr_scheduler.state_dict()
    os.makedirs(save_path, exist_ok=True)
    path = os.path.join(save_path, filename)
    torch.save(checkpoint, path)
    return path


def load_checkpoint(model_path, device, model=None, optimizer=None, scheduler=None):
    """utils/nbody_utils.py:1379-1413 (and trainer.py:72-86): returns the model; ``load_checkpoint.last`` holds
    ``step_count`` / ``best_metrics`` of the file for the trainer's resume path."""
    checkpoint = torch.load(model_path, map_location=device, weights_only=False)
    info = {"step_count": None, "best_metrics": None}
    if isinstance(checkpoint, dict) and "model_state_dict" in checkpoint:
        if model is not None:
            model.load_state_dict(checkpoint["model_state_dict"])
        if optimizer is not None and checkpoint.get("optimizer_state_dict"):
            optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
        if scheduler is not None and "scheduler_state_dict" in checkpoint:
            scheduler.load_state_dict(checkpoint["scheduler_state_dict"])
        info = {"step_count": checkpoint.get("step_count"), "best_metrics": checkpoint.get("best_metrics")}
    elif model is not None:  # a bare state_dict
        model.load_state_dict(checkpoint)
    else:
        raise ValueError("Checkpoint does not contain a dictionary and no model provided.")
    load_checkpoint.last = info
    return model


load_checkpoint.last = {"step_count": None, "best_metrics": None}


def load_model_for_inference(model_path, model_type, device, **model_kwargs):
    """utils/nbody_utils.py:1316-1376 for model_type='segnn': default-constructed SEGNN (the reference calls
    ``SEGNN()``; pass the training-time ``hidden_features`` / ``num_layers`` / ``lmax_h`` as keyword arguments for any
    other size), weights from the checkpoint, ``.eval()`` on ``device``."""
    if model_type != "segnn":
        raise ValueError(f"Unsupported model_type: {model_type}")
    from .segnn import SEGNN
    model = SEGNN(**model_kwargs)
    load_checkpoint(model_path, "cpu", model=model)
    return model.eval().to(device)


def save_run_metadata(save_dir_path: str, model, dataset, dataset_name: Optional[str] = None) -> None:
    """trainer.py:526-537: ``model_params.json`` and ``<dataset_name>_dataset/metadata.json``."""
    os.makedirs(save_dir_path, exist_ok=True)
    with open(os.path.join(save_dir_path, "model_params.json"), "w") as f:
        json.dump(model.get_serializable_attributes(), f, indent=4)
    name = dataset_name if dataset_name is not None else dataset.dataset_name
    ds_dir = os.path.join(save_dir_path, f"{name}_dataset")
    os.makedirs(ds_dir, exist_ok=True)
    with open(os.path.join(ds_dir, "metadata.json"), "w") as f:
        json.dump(dataset.get_serializable_attributes(), f, indent=4)


def get_dataset_metadata_path(path: str) -> str:
    """utils/nbody_utils.py:1446-1488: the outermost ancestor named like a run directory
    (``YYYY-MM-DD_HH-MM-SS...``) holds ``nbody_small_dataset/metadata.json``."""
    current = os.path.abspath(path)
    pattern = re.compile(r"\d{4}-\d{2}-\d{2}_\d{2}-\d{2}-\d{2}")
    run_root = None
    while True:
        if pattern.match(os.path.basename(current)):
            run_root = current
        parent = os.path.dirname(current)
        if parent == current:
            break
        current = parent
    if run_root is None:
        raise FileNotFoundError("Run directory root not found")
    return os.path.join(run_root, "nbody_small_dataset", "metadata.json")


def load_dataset_from_metadata_file(metadata_file_path: str, n_bodies=None, **overrides):
    """visualization_utils.py:1438-1452: only the keys that are constructor parameters are passed on, so ``n_balls``
    is dropped and the number of bodies falls back to the default 5 unless ``n_bodies`` is given -- exactly the
    reference's behaviour."""
    from .dataloader import GravityDatasetOtf
    with open(metadata_file_path, "r") as f:
        metadata = json.load(f)
    expected = [p for p in inspect.signature(GravityDatasetOtf.__init__).parameters if p not in ("self", "_ignored")]
    filtered = {k: metadata[k] for k in expected if k in metadata}
    if n_bodies is not None:
        filtered["num_nodes"] = n_bodies
    filtered.update(overrides)
    return GravityDatasetOtf(**filtered)
