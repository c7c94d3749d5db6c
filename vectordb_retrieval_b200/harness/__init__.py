"""Minimal experiment / benchmark harness speaking the reference's YAML schema and result keys
(src/experiments/*, src/benchmark/runner.py) - the *caller* of the hot path, written fresh so
``scripts/run_full_benchmark.py --config <reference-style yaml>`` runs without faiss / matplotlib."""
from .config import ExperimentConfig
from .experiment_runner import ExperimentRunner
from .runner import BenchmarkRunner

__all__ = ["ExperimentConfig", "ExperimentRunner", "BenchmarkRunner"]
