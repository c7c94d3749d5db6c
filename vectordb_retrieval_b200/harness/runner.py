"""Multi-dataset benchmark driver speaking the reference's YAML schema
(reference: src/benchmark/runner.py:22-215, 232-299).

Top-level keys: ``indexers``, ``searchers``, ``algorithms``, ``datasets[]``, ``output_dir``,
``data_dir``, ``n_queries``, ``topk``, ``repeat``, ``seed``, ``query_batch_size``.  A dataset entry is
a name or ``{name, metric, n_queries, topk, dataset_options, algorithms: {overrides}}``; the
dataset metric overwrites every algorithm's metric; ``indexer_ref`` / ``searcher_ref`` resolve
against the global tables with inline overrides deep-merged on top and imply ``type: Composite``.
A failing dataset is logged and skipped.  Reports: ``all_results.json`` and
``benchmark_summary.md`` (the SVG one-pager of the reference is cosmetic and omitted)."""
from __future__ import annotations

import copy
import datetime
import json
import logging
import os
import time
from typing import Any, Dict, Optional, Tuple

import yaml

from ..algorithms import get_algorithm_instance
from .config import ExperimentConfig
from .experiment_runner import ExperimentRunner


class BenchmarkRunner:
    def __init__(self, config_file: str, output_dir: str = "benchmark_results") -> None:
        self.config_file = config_file
        self.timestamp = datetime.datetime.now().strftime("%Y%m%d_%H%M%S")
        with open(config_file, "r") as f:
            self.config = json.load(f) if config_file.endswith(".json") else yaml.safe_load(f)
        self.global_indexers = copy.deepcopy(self.config.get("indexers", {}))
        self.global_searchers = copy.deepcopy(self.config.get("searchers", {}))
        self.output_dir = os.path.join(self.config.get("output_dir", output_dir), f"benchmark_{self.timestamp}")
        os.makedirs(self.output_dir, exist_ok=True)
        self.logger = logging.getLogger("benchmark")
        self.logger.setLevel(logging.INFO)
        if not self.logger.handlers:
            self.logger.addHandler(logging.StreamHandler())
        self.all_results: Dict[str, Any] = {}

    @staticmethod
    def _normalize_dataset_entry(entry: Any) -> Tuple[str, Dict[str, Any]]:
        if isinstance(entry, str):
            return entry, {}
        if isinstance(entry, dict):
            if "name" not in entry:
                raise ValueError("Dataset configuration entries must include a 'name' key")
            return entry["name"], {k: v for k, v in entry.items() if k != "name"}
        raise ValueError("Dataset configuration entries must be strings or dictionaries")

    @staticmethod
    def _deep_merge_dict(base: Dict[str, Any], updates: Dict[str, Any]) -> Dict[str, Any]:
        out = copy.deepcopy(base)
        for key, value in updates.items():
            if isinstance(out.get(key), dict) and isinstance(value, dict):
                out[key] = BenchmarkRunner._deep_merge_dict(out[key], value)
            else:
                out[key] = copy.deepcopy(value)
        return out

    def _materialize_component(self, ref_name: Optional[str], inline_cfg: Any, registry: Dict[str, Dict[str, Any]],
                               label: str) -> Optional[Dict[str, Any]]:
        def lookup(ref: str) -> Dict[str, Any]:
            if ref not in registry:
                raise ValueError(f"Unknown {label} reference '{ref}'. Available: {list(registry.keys())}")
            return copy.deepcopy(registry[ref])

        config = lookup(ref_name) if ref_name else None
        if inline_cfg is None:
            return config
        if isinstance(inline_cfg, str):
            inline = lookup(inline_cfg)
        elif isinstance(inline_cfg, dict):
            inline = copy.deepcopy(inline_cfg)
        else:
            raise TypeError(f"{label.capitalize()} configuration must be a dict or string reference, got {type(inline_cfg)}")
        return inline if config is None else self._deep_merge_dict(config, inline)

    def _resolve_modular_components(self, cfg: Dict[str, Any]) -> None:
        indexer = self._materialize_component(cfg.pop("indexer_ref", None), cfg.get("indexer"), self.global_indexers, "indexer")
        searcher = self._materialize_component(cfg.pop("searcher_ref", None), cfg.get("searcher"), self.global_searchers, "searcher")
        if indexer is not None:
            cfg["indexer"] = indexer
        if searcher is not None:
            cfg["searcher"] = searcher
        if indexer is not None or searcher is not None:
            cfg.setdefault("type", "Composite")

    def _algorithms_for(self, overrides: Dict[str, Any], metric: Optional[str]) -> Dict[str, Dict[str, Any]]:
        out: Dict[str, Dict[str, Any]] = {}
        for name, base in copy.deepcopy(self.config.get("algorithms", {})).items():
            merged = copy.deepcopy(base)
            merged.update(copy.deepcopy(overrides.get(name, {})))
            if metric is not None:
                merged["metric"] = metric
            self._resolve_modular_components(merged)
            out[name] = merged
        for name, cfg in overrides.items():
            if name not in out:
                merged = copy.deepcopy(cfg)
                if metric is not None and "metric" not in merged:
                    merged["metric"] = metric
                self._resolve_modular_components(merged)
                out[name] = merged
        return out

    def run(self) -> Dict[str, Any]:
        start = time.time()
        for entry in self.config.get("datasets", ["random"]):
            name, opts = self._normalize_dataset_entry(entry)
            metric = opts.get("metric")
            kwargs: Dict[str, Any] = dict(
                dataset=name,
                data_dir=opts.get("data_dir", self.config.get("data_dir", "data")),
                force_download=self.config.get("force_download", False),
                n_queries=opts.get("n_queries", self.config.get("n_queries", 1000)),
                topk=opts.get("topk", self.config.get("topk", 100)),
                repeat=opts.get("repeat", self.config.get("repeat", 1)),
                algorithms=self._algorithms_for(opts.get("algorithms", {}), metric),
                seed=opts.get("seed", self.config.get("seed", 42)),
                output_prefix=opts.get("output_prefix", f"{name}_{self.timestamp}"),
            )
            qbs = opts.get("query_batch_size", self.config.get("query_batch_size"))
            if qbs is not None:
                kwargs["query_batch_size"] = qbs
            if metric is not None:
                kwargs["metric"] = metric
            ds_opts = copy.deepcopy(opts.get("dataset_options") or opts.get("options") or {})
            if metric is not None:
                ds_opts.setdefault("metric", metric)
            if ds_opts:
                kwargs["dataset_options"] = ds_opts
            cfg = ExperimentConfig(**kwargs)
            out_dir = os.path.join(self.output_dir, name)
            os.makedirs(out_dir, exist_ok=True)
            cfg.save(os.path.join(out_dir, f"{name}_config.yaml"))
            runner = ExperimentRunner(cfg, output_dir=out_dir)
            try:
                runner.load_dataset()
                dimension = runner.dataset.train_vectors.shape[1]
                for alg_name, alg_cfg in cfg.algorithms.items():
                    alg_cfg = copy.deepcopy(alg_cfg)
                    alg_type = alg_cfg.pop("type")
                    runner.register_algorithm(get_algorithm_instance(alg_type, dimension, name=alg_name, **alg_cfg), name=alg_name)
                results = runner.run()
                self.all_results[name] = results
                with open(os.path.join(out_dir, f"{name}_results.json"), "w") as f:
                    json.dump(results, f, indent=2, default=str)
            except Exception as exc:      # noqa: BLE001 - a failing dataset is logged and skipped (runner.py:197-198)
                self.logger.error(f"Error running experiments for dataset {name}: {exc}", exc_info=True)
        with open(os.path.join(self.output_dir, "all_results.json"), "w") as f:
            json.dump(self.all_results, f, indent=2, default=str)
        self._write_summary()
        self.logger.info(f"Benchmark completed in {time.time() - start:.2f} seconds")
        return self.all_results

    def _write_summary(self) -> None:
        lines = ["# Benchmark summary", "", f"config: `{self.config_file}`", ""]
        for ds, results in self.all_results.items():
            lines += [f"## {ds}", "", "| algorithm | recall | qps | mean query ms | build s | index MB |", "|---|---|---|---|---|---|"]
            for name, r in results.items():
                lines.append(f"| {name} | {r.get('recall', float('nan')):.4f} | {r.get('qps', 0.0):.2f} | "
                             f"{r.get('mean_query_time_ms', 0.0):.4f} | {r.get('build_time_s', 0.0):.3f} | "
                             f"{r.get('index_memory_mb', 0.0):.1f} |")
            lines.append("")
        with open(os.path.join(self.output_dir, "benchmark_summary.md"), "w") as f:
            f.write("\n".join(lines))
