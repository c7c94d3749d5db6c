"""Per-dataset experiment loop: the caller of the hot path
(reference: src/experiments/experiment_runner.py:28-133, 259-488).

Kept from the reference because results must be comparable: ``np.random.seed(config.seed)`` then
``np.random.choice`` for the query subset, wall-clock ``time.time()`` around ``build_index`` and
around every ``batch_search`` chunk of ``query_batch_size`` queries, only the *indices* of a
result are consumed, the batch API falling back to per-query ``search`` on the same four
exception types, and the metric key names of the result JSON.  Added: one untimed warm-up
``batch_search`` of a single query when ``warmup`` is set (the reference times the very first
call, which on a GPU would charge context creation to the first batch), and a ``roofline`` block
when the algorithm exposes ``last_kernel_ms``.  The per-algorithm ``persistence`` block
(``enabled, mode {build_only, retrieve_only, build_and_retrieve}, artifact_dir, path_policy
{fixed, versioned}, version_tag, fail_if_missing``; reference :163-186, 242-372) drives the
``save_index / load_index`` protocol.  Plots are out of scope."""
from __future__ import annotations

import copy
import hashlib
import json
import logging
import os
import time
from datetime import datetime
from typing import Any, Dict, Optional, Tuple

import numpy as np
import yaml

from ..algorithms.base_algorithm import BaseAlgorithm
from . import metrics as M
from .config import ExperimentConfig
from .dataset import Dataset


class ExperimentRunner:
    def __init__(self, config: ExperimentConfig, output_dir: str = "results", warmup: bool = True) -> None:
        self.config = config
        self.output_dir = output_dir
        self.warmup = warmup
        self.dataset: Optional[Dataset] = None
        self.algorithms: Dict[str, BaseAlgorithm] = {}
        self.results: Dict[str, Dict[str, Any]] = {}
        os.makedirs(self.output_dir, exist_ok=True)
        self.logger = logging.getLogger("experiment_runner")

    def load_dataset(self) -> None:
        if self.dataset is None:
            ds = Dataset(self.config.dataset, self.config.data_dir, options=self.config.dataset_options)
            ds.load(force_download=self.config.force_download)
            self.dataset = ds

    def register_algorithm(self, algorithm: BaseAlgorithm, name: Optional[str] = None) -> None:
        if not isinstance(algorithm, BaseAlgorithm):
            raise TypeError("algorithm must inherit from BaseAlgorithm")
        algorithm_name = name or algorithm.get_name()
        if not algorithm_name:
            raise ValueError("Algorithm name must be provided")
        algorithm.name = algorithm_name
        self.algorithms[algorithm_name] = algorithm

    def _select_query_subset(self, queries, ground_truth):
        n_available = len(queries)
        target = min(self.config.n_queries, n_available) if self.config.n_queries else n_available
        if target >= n_available:
            return queries, ground_truth
        pick = np.random.choice(n_available, target, replace=False)
        return queries[pick], (ground_truth[pick] if ground_truth is not None else None)

    @staticmethod
    def _indices_of(result: Any, rows: int, k: int) -> np.ndarray:
        if isinstance(result, tuple):
            if len(result) != 2:
                raise ValueError("batch_search must return (distances, indices)")
            result = result[1]
        arr = np.asarray(result)
        if arr.ndim == 1:
            arr = arr.reshape(1, -1)
        if arr.ndim != 2 or arr.shape[0] != rows:
            raise ValueError(f"batch_search returned shape {arr.shape}, expected ({rows}, {k})")
        if arr.shape[1] < k:
            padded = np.full((rows, k), -1, dtype=np.int64)
            padded[:, : arr.shape[1]] = arr
            arr = padded
        return arr[:, :k].astype(np.int64, copy=False)

    def _memory_mb(self, algorithm: BaseAlgorithm, train: np.ndarray) -> float:
        fn = getattr(algorithm, "get_memory_usage", None)
        if callable(fn):
            try:
                return float(fn()) / (1024.0 * 1024.0)
            except Exception:          # noqa: BLE001 - estimate only
                pass
        return float(train.shape[0]) * train.shape[1] * 4 / (1024.0 * 1024.0)

    # ---- persistence (reference: experiment_runner.py:163-186, 242-257, 267-372) -----------------
    def _persistence_cfg(self, name: str) -> Dict[str, Any]:
        raw = self.config.algorithms.get(name, {})
        raw = raw.get("persistence", {}) if isinstance(raw, dict) else {}
        if not isinstance(raw, dict):
            return {}
        cfg = copy.deepcopy(raw)
        cfg["mode"] = str(cfg.get("mode", "build_and_retrieve")).strip().lower()
        if cfg["mode"] not in ("build_only", "retrieve_only", "build_and_retrieve"):
            raise ValueError(f"Unsupported persistence mode '{cfg['mode']}'")
        cfg["path_policy"] = str(cfg.get("path_policy", "fixed")).strip().lower()
        if cfg["path_policy"] not in ("fixed", "versioned"):
            raise ValueError(f"Unsupported persistence path_policy '{cfg['path_policy']}'")
        cfg["enabled"] = bool(cfg.get("enabled", False))
        return cfg

    @staticmethod
    def _stable_hash(payload: Any) -> str:
        return hashlib.sha256(json.dumps(payload, sort_keys=True, default=str).encode()).hexdigest()[:16]

    def _persistence_context(self, name: str, train: np.ndarray) -> Dict[str, Any]:
        algo_cfg = copy.deepcopy(self.config.algorithms.get(name, {}))
        if isinstance(algo_cfg, dict):
            algo_cfg.pop("persistence", None)
        fp_payload = {"dataset": self.config.dataset, "dataset_options": self.config.dataset_options,
                      "n_train": int(train.shape[0]), "dimensions": int(train.shape[1]), "dtype": str(train.dtype)}
        return {"dataset_fingerprint": self._stable_hash(fp_payload), "dataset_fingerprint_payload": fp_payload,
                "config_hash": self._stable_hash({"algorithm_name": name, "algorithm_config": algo_cfg, **fp_payload}),
                "force_rebuild": False}

    def _build_or_load(self, name: str, algorithm: BaseAlgorithm, train: np.ndarray) -> Dict[str, Any]:
        pcfg = self._persistence_cfg(name)
        info: Dict[str, Any] = {"build_time_s": 0.0, "index_load_time_s": 0.0, "index_source": "built"}
        if not pcfg.get("enabled"):
            t0 = time.time()
            algorithm.build_index(train)
            info["build_time_s"] = time.time() - t0
            return info
        ctx = self._persistence_context(name, train)
        ctx["force_rebuild"] = bool(pcfg.get("force_rebuild", False))
        if not pcfg.get("artifact_dir"):
            raise ValueError(f"Algorithm '{name}' has persistence enabled but no persistence.artifact_dir configured.")
        pdir = str(pcfg["artifact_dir"])
        if pcfg["path_policy"] == "versioned":
            tag = pcfg.get("version_tag")
            pdir = os.path.join(pdir, str(tag).strip() if tag else ctx["dataset_fingerprint"])
        info.update(persistence_mode=pcfg["mode"], persist_dir=pdir, dataset_fingerprint=ctx["dataset_fingerprint"],
                    config_hash=ctx["config_hash"])
        if pcfg["mode"] == "retrieve_only" and os.path.isdir(pdir):
            t0 = time.time()
            loaded = algorithm.load_index(pdir, context=ctx)
            info["index_load_time_s"] = time.time() - t0
            info["index_source"] = "loaded"
            info["build_time_s"] = float(loaded.get("build_time_s", 0.0) or 0.0)
            return info
        if pcfg["mode"] == "retrieve_only" and pcfg.get("fail_if_missing", True):
            raise FileNotFoundError(f"Missing persisted index for '{name}' at {pdir}. "
                                    "Run build_only (or build_and_retrieve) first to create the artifact.")
        t0 = time.time()
        algorithm.build_index(train)
        info["build_time_s"] = time.time() - t0
        if pcfg["mode"] in ("build_only", "build_and_retrieve"):
            ctx["build_metrics"] = {"build_time_s": float(info["build_time_s"]), "n_train": int(train.shape[0]),
                                    "dimensions": int(train.shape[1]), "timestamp": datetime.now().isoformat()}
            algorithm.save_index(pdir, context=ctx)
        return info

    def _run_single_algorithm(self, name: str, algorithm: BaseAlgorithm, train: np.ndarray, queries: np.ndarray
                              ) -> Tuple[Dict[str, Any], Optional[np.ndarray], Optional[np.ndarray]]:
        pinfo = self._build_or_load(name, algorithm, train)
        build_time = pinfo["build_time_s"]
        if pinfo.get("persistence_mode") == "build_only":
            out = {"algorithm": name, "parameters": algorithm.get_parameters(), "dataset": self.config.dataset,
                   "n_train": int(train.shape[0]), "n_test": int(len(queries)), "dimensions": int(train.shape[1]),
                   "topk": self.config.topk, "index_memory_mb": self._memory_mb(algorithm, train), "qps": 0.0,
                   "mean_query_time_ms": 0.0, "total_query_time_s": 0.0, "status": "build_only",
                   "timestamp": datetime.now().isoformat(), **pinfo}
            return out, None, None
        k = self.config.topk
        nq = len(queries)
        indices = np.full((nq, k), -1, dtype=np.int64)
        query_times = np.zeros(nq, dtype=float)
        total = 0.0
        used_batch = False
        if nq > 0:
            cfg_bs = max(int(getattr(self.config, "query_batch_size", 0) or 0), 0)
            bs = nq if cfg_bs == 0 else min(cfg_bs, nq)
            try:
                if self.warmup:
                    algorithm.batch_search(queries[:1], k=k)
                cursor = 0
                while cursor < nq:
                    end = min(cursor + bs, nq)
                    t = time.time()
                    res = algorithm.batch_search(queries[cursor:end], k=k)
                    elapsed = time.time() - t
                    indices[cursor:end] = self._indices_of(res, end - cursor, k)
                    query_times[cursor:end] = elapsed / max(end - cursor, 1)
                    total += elapsed
                    cursor = end
                used_batch = True
            except (AttributeError, NotImplementedError, TypeError, ValueError):
                indices.fill(-1)
                query_times.fill(0.0)
                total = 0.0
        if not used_batch:
            for i, q in enumerate(queries):
                t = time.time()
                _, single = algorithm.search(q, k=k)
                query_times[i] = time.time() - t
                indices[i] = single
                total += query_times[i]
        total = max(total, float(query_times.sum()))
        out: Dict[str, Any] = {
            "algorithm": name,
            "parameters": algorithm.get_parameters(),
            "dataset": self.config.dataset,
            "n_train": int(train.shape[0]),
            "n_test": int(nq),
            "dimensions": int(train.shape[1]),
            "topk": k,
            "build_time_s": float(build_time),
            "index_memory_mb": self._memory_mb(algorithm, train),
            "qps": float(nq / total) if total > 0 else 0.0,
            "mean_query_time_ms": float(total / max(nq, 1) * 1000.0),
            "total_query_time_s": float(total),
            "timestamp": datetime.now().isoformat(),
        }
        out.update({key: v for key, v in pinfo.items() if key != "build_time_s"})
        ops = algorithm.get_operations()
        if ops:
            out["operations"] = ops
        return out, indices, query_times

    def run(self) -> Dict[str, Dict[str, Any]]:
        if not self.algorithms:
            raise RuntimeError("No algorithms registered for the experiment")
        self.load_dataset()
        np.random.seed(self.config.seed)
        train, test, gt = self.dataset.train_vectors, self.dataset.test_vectors, self.dataset.ground_truth
        if train is None or test is None:
            raise RuntimeError("Dataset did not provide train/test vectors")
        test, gt = self._select_query_subset(test, gt)
        self.results = {}
        outputs: Dict[str, Tuple[np.ndarray, np.ndarray]] = {}
        for name, algorithm in self.algorithms.items():
            metrics, idx, times = self._run_single_algorithm(name, algorithm, train, test)
            self.results[name] = metrics
            if idx is not None:
                outputs[name] = (idx, times)
        if gt is not None:
            for name, (idx, times) in outputs.items():
                ev = M.evaluate(gt, idx, times)
                self.results[name].update(ev)
                key = f"recall@{min(100, self.config.topk)}"
                if key in ev:
                    self.results[name]["recall"] = ev[key]
                else:
                    rec = sorted((m for m in ev if m.startswith("recall@")), key=lambda m: int(m.split("@")[-1]))
                    if rec:
                        self.results[name]["recall"] = ev[rec[-1]]
        for name, res in self.results.items():
            with open(os.path.join(self.output_dir, f"{name}_results.json"), "w") as f:
                json.dump(res, f, indent=2, default=str)
        prefix = self.config.output_prefix
        with open(os.path.join(self.output_dir, f"{prefix}_all_results.json"), "w") as f:
            json.dump(self.results, f, indent=2, default=str)
        with open(os.path.join(self.output_dir, f"{prefix}_config.yaml"), "w") as f:
            yaml.dump(self.config.to_dict(), f)
        return self.results
