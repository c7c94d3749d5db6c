"""Flat kwargs experiment configuration with YAML I/O (reference: src/experiments/config.py:5-101).
Unknown keys are ignored; a dataset-wide ``metric`` becomes each algorithm's default."""
from __future__ import annotations

import copy
from typing import Any, Dict

import yaml

_DEFAULT_ALGORITHMS = {"exact": {"type": "ExactSearch", "metric": "l2"}}


class ExperimentConfig:
    def __init__(self, **kwargs: Any) -> None:
        self.dataset = kwargs.get("dataset", "random")
        self.data_dir = kwargs.get("data_dir", "data")
        self.force_download = kwargs.get("force_download", False)
        self.dataset_options = copy.deepcopy(kwargs.get("dataset_options", {}))
        self.n_queries = kwargs.get("n_queries", 1000)
        self.topk = kwargs.get("topk", 100)
        self.repeat = kwargs.get("repeat", 1)               # parsed, unused - as in the reference
        self.query_batch_size = kwargs.get("query_batch_size", 0)   # 0 = all queries in one call
        self.algorithms = copy.deepcopy(kwargs.get("algorithms", _DEFAULT_ALGORITHMS))
        self.metric = kwargs.get("metric")
        if self.metric is not None:
            for cfg in self.algorithms.values():
                if isinstance(cfg, dict):
                    cfg.setdefault("metric", self.metric)
        self.seed = kwargs.get("seed", 42)
        self.output_prefix = kwargs.get("output_prefix", "experiment")

    @classmethod
    def from_yaml(cls, yaml_file: str) -> "ExperimentConfig":
        with open(yaml_file, "r") as f:
            return cls(**yaml.safe_load(f))

    def to_dict(self) -> Dict[str, Any]:
        out = {k: getattr(self, k) for k in ("dataset", "data_dir", "force_download", "dataset_options", "n_queries", "topk",
                                             "repeat", "query_batch_size", "algorithms", "seed", "output_prefix")}
        if self.metric is not None:
            out["metric"] = self.metric
        return out

    def save(self, output_file: str) -> None:
        with open(output_file, "w") as f:
            yaml.dump(self.to_dict(), f)

    def __str__(self) -> str:
        return yaml.dump(self.to_dict())
