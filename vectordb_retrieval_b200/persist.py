"""Index persistence behind the reference's optional ``save_index / load_index`` protocol
(src/algorithms/base_algorithm.py:98-120, driven by src/experiments/experiment_runner.py:308-344;
in the reference only the cover tree implements it, src/algorithms/covertree_v2_2.py:101-285).

Layout of an artifact directory, modelled on the cover tree's: one ``.npy`` per device array (the
HBM layouts of DESIGN.md section 2, copied bit for bit), ``manifest.json`` (format version, index
kind, shapes, the caller's context: dataset fingerprint, config hash, build metrics) and an empty
``WRITE_COMPLETE`` sentinel written last - a directory without it is treated as missing."""
from __future__ import annotations

import json
import os
from typing import Any, Dict, Optional

import numpy as np

FORMAT_VERSION = 1
SENTINEL = "WRITE_COMPLETE"


def write_artifact(artifact_dir: str, kind: str, arrays: Dict[str, np.ndarray], meta: Dict[str, Any],
                   context: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
    os.makedirs(artifact_dir, exist_ok=True)
    sentinel = os.path.join(artifact_dir, SENTINEL)
    if os.path.exists(sentinel):
        os.remove(sentinel)
    files, total = [], 0
    for name, arr in arrays.items():
        path = os.path.join(artifact_dir, f"{name}.npy")
        np.save(path, arr)
        files.append(f"{name}.npy")
        total += os.path.getsize(path)
    ctx = context or {}
    manifest = {"format_version": FORMAT_VERSION, "kind": kind, "meta": meta, "files": files,
                "dataset_fingerprint": ctx.get("dataset_fingerprint"), "config_hash": ctx.get("config_hash"),
                "build_metrics": ctx.get("build_metrics", {})}
    with open(os.path.join(artifact_dir, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=2, default=str)
    open(sentinel, "w").close()
    return {"artifact_dir": artifact_dir, "files": files, "bytes": int(total), "kind": kind}


def read_artifact(artifact_dir: str, kind: str, context: Optional[Dict[str, Any]] = None):
    """-> (arrays, manifest).  FileNotFoundError if absent or incomplete; RuntimeError on a kind /
    version / fingerprint mismatch (a stale artifact must not be served silently)."""
    if not os.path.isdir(artifact_dir) or not os.path.exists(os.path.join(artifact_dir, SENTINEL)):
        raise FileNotFoundError(f"no complete index artifact at {artifact_dir}")
    with open(os.path.join(artifact_dir, "manifest.json")) as f:
        manifest = json.load(f)
    if manifest.get("format_version") != FORMAT_VERSION or manifest.get("kind") != kind:
        raise RuntimeError(f"artifact at {artifact_dir} is '{manifest.get('kind')}' v{manifest.get('format_version')}, "
                           f"expected '{kind}' v{FORMAT_VERSION}")
    ctx = context or {}
    want = ctx.get("dataset_fingerprint")
    if want is not None and manifest.get("dataset_fingerprint") not in (None, want):
        raise RuntimeError(f"artifact at {artifact_dir} was built for another dataset (fingerprint mismatch)")
    arrays = {name[:-4]: np.load(os.path.join(artifact_dir, name)) for name in manifest["files"]}
    return arrays, manifest
