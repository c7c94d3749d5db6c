"""Retrieval metrics the harness reports (reference: src/benchmark/metrics.py:4-34 and the
Evaluator keys of src/benchmark/evaluation.py:31-66)."""
from __future__ import annotations

from typing import Any, Dict, Sequence

import numpy as np


def recall_at_k(ground_truth: np.ndarray, predicted: np.ndarray, k: int) -> float:
    """Mean over queries of |gt[:k] & pred[:k]| / |gt[:k]| (k clipped to the predicted width)."""
    k = min(k, predicted.shape[1])
    total = 0.0
    for gt_row, pr_row in zip(ground_truth, predicted):
        gt = set(gt_row[:k].tolist())
        total += (len(gt & set(pr_row[:k].tolist())) / len(gt)) if gt else 0.0
    return total / max(len(ground_truth), 1)


def precision_at_k(ground_truth: np.ndarray, predicted: np.ndarray, k: int) -> float:
    k = min(k, predicted.shape[1])
    total = 0.0
    for gt_row, pr_row in zip(ground_truth, predicted):
        total += len(set(gt_row[:k].tolist()) & set(pr_row[:k].tolist())) / k
    return total / max(len(ground_truth), 1)


def mean_average_precision(ground_truth: np.ndarray, predicted: np.ndarray, k: int) -> float:
    k = min(k, predicted.shape[1])
    total = 0.0
    for gt_row, pr_row in zip(ground_truth, predicted):
        gt = set(gt_row[:k].tolist())
        hits, score = 0, 0.0
        for rank, p in enumerate(pr_row[:k].tolist(), start=1):
            if p in gt:
                hits += 1
                score += hits / rank
        total += score / min(len(gt), k) if gt else 0.0
    return total / max(len(ground_truth), 1)


def evaluate(ground_truth: np.ndarray, predicted: np.ndarray, query_times: np.ndarray,
             k_values: Sequence[int] = (1, 10, 100)) -> Dict[str, Any]:
    """Evaluator.evaluate: recall/precision at k in {1, 10, 100} that fit, map@10, and qps as
    1 / mean(per-query time) (the reference overwrites the loop's qps with this, evaluation.py:57)."""
    out: Dict[str, Any] = {}
    for k in k_values:
        if k <= predicted.shape[1]:
            out[f"recall@{k}"] = recall_at_k(ground_truth, predicted, k)
            out[f"precision@{k}"] = precision_at_k(ground_truth, predicted, k)
    if predicted.shape[1] >= 10:
        out["map@10"] = mean_average_precision(ground_truth, predicted, 10)
    out["qps"] = float(1.0 / np.mean(query_times))
    out["mean_query_time"] = float(np.mean(query_times) * 1000)
    out["median_query_time"] = float(np.median(query_times) * 1000)
    out["min_query_time"] = float(np.min(query_times) * 1000)
    out["max_query_time"] = float(np.max(query_times) * 1000)
    return out
