"""Experiment configuration with YAML I/O, table-driven (reference: src/experiments/config.py:5-101).
Same keys, defaults and behaviour: unknown keys are ignored, a dataset-wide ``metric`` becomes each
algorithm's default, ``metric`` is only serialised when set."""
from __future__ import annotations

import copy
from typing import Any, Dict

import yaml

# key -> default, in the order the reference serialises them; mutable defaults are deep-copied per instance
_FIELDS: Dict[str, Any] = {
    "dataset": "random",
    "data_dir": "data",
    "force_download": False,
    "dataset_options": {},
    "n_queries": 1000,
    "topk": 100,
    "repeat": 1,                 # parsed, unused - as in the reference
    "query_batch_size": 0,       # 0 = all queries in one call
    "algorithms": {"exact": {"type": "ExactSearch", "metric": "l2"}},
    "seed": 42,
    "output_prefix": "experiment",
}


class ExperimentConfig:
    def __init__(self, **kwargs: Any) -> None:
        for key, default in _FIELDS.items():
            value = kwargs.get(key, default)
            setattr(self, key, copy.deepcopy(value) if isinstance(value, (dict, list)) else value)
        self.metric = kwargs.get("metric")
        if self.metric is not None:
            for cfg in self.algorithms.values():
                if isinstance(cfg, dict):
                    cfg.setdefault("metric", self.metric)

    @classmethod
    def from_yaml(cls, yaml_file: str) -> "ExperimentConfig":
        with open(yaml_file, "r") as f:
            return cls(**yaml.safe_load(f))

    def to_dict(self) -> Dict[str, Any]:
        out = {key: getattr(self, key) for key in _FIELDS}
        if self.metric is not None:
            out["metric"] = self.metric
        return out

    def save(self, output_file: str) -> None:
        with open(output_file, "w") as f:
            yaml.dump(self.to_dict(), f)

    def __str__(self) -> str:
        return yaml.dump(self.to_dict())
