"""Datasets for the harness: the reference's ``random`` generator restated
(src/benchmark/dataset.py:473-504: ``np.random.seed(seed)``, ``randn`` base then queries, L2
ground truth by argsort), plus synthetic generators of the BASELINE shapes and ``.npy`` /
``.fvecs`` readers (dataset.py:522-574).  Downloaders and text pipelines are out of scope.

Ground truth: small problems use the reference's NumPy recipe (independent of the kernels under
test); large ones use the exact GPU search and say so in ``ground_truth_source``."""
from __future__ import annotations

import os
from typing import Any, Dict, Optional

import numpy as np

_CPU_GT_LIMIT = 2.0e8      # n_train * n_test above which NumPy ground truth is too slow


def read_fvecs(path: str, limit: Optional[int] = None) -> np.ndarray:
    raw = np.fromfile(path, dtype=np.int32)
    d = int(raw[0])
    out = raw.reshape(-1, d + 1)[:, 1:].view(np.float32)
    return np.ascontiguousarray(out[:limit] if limit else out)


def read_ivecs(path: str, limit: Optional[int] = None) -> np.ndarray:
    raw = np.fromfile(path, dtype=np.int32)
    d = int(raw[0])
    out = raw.reshape(-1, d + 1)[:, 1:]
    return np.ascontiguousarray(out[:limit] if limit else out)


class Dataset:
    def __init__(self, name: str, data_dir: str = "data", options: Optional[Dict[str, Any]] = None) -> None:
        self.name = name
        self.data_dir = data_dir
        self.options = dict(options or {})
        self.train_vectors: Optional[np.ndarray] = None
        self.test_vectors: Optional[np.ndarray] = None
        self.ground_truth: Optional[np.ndarray] = None
        self.ground_truth_source = "none"

    # ------------------------------------------------------------------ generators
    def _random(self, dims: int, train: int, test: int) -> None:
        np.random.seed(self.options.get("seed", 42))
        self.train_vectors = np.random.randn(train, dims).astype(np.float32)
        self.test_vectors = np.random.randn(test, dims).astype(np.float32)

    def _clustered(self, dims: int, train: int, test: int, centers: int, sigma: float) -> None:
        """N(0,1) centres + N(0, sigma) noise: gives IVF / LSH a non-trivial recall curve (config C3)."""
        rng = np.random.RandomState(self.options.get("seed", 42))
        mu = rng.randn(centers, dims).astype(np.float32)
        self.train_vectors = (mu[rng.randint(0, centers, train)] + sigma * rng.randn(train, dims)).astype(np.float32)
        self.test_vectors = (mu[rng.randint(0, centers, test)] + sigma * rng.randn(test, dims)).astype(np.float32)

    def load(self, force_download: bool = False) -> None:
        o = self.options
        gt_k = int(o.get("ground_truth_k", 100))
        metric = o.get("metric", "l2")
        if self.name == "random":
            self._random(int(o.get("dimensions", 128)), int(o.get("train_size", 10_000)), int(o.get("test_size", 1_000)))
        elif self.name in ("sift1m_shape", "synthetic"):
            self._random(int(o.get("dimensions", 128)), int(o.get("train_size", 1_000_000)), int(o.get("test_size", 10_000)))
        elif self.name == "glove50_shape":
            self._clustered(int(o.get("dimensions", 50)), int(o.get("train_size", 1_200_000)), int(o.get("test_size", 10_000)),
                            int(o.get("centers", 64)), float(o.get("sigma", 0.3)))
        elif self.name == "npy":
            self.train_vectors = np.load(o["train_path"], mmap_mode="r")
            self.test_vectors = np.load(o["test_path"], mmap_mode="r")
            if o.get("ground_truth_path"):
                self.ground_truth = np.load(o["ground_truth_path"])
                self.ground_truth_source = "file"
        elif self.name in ("sift1m", "fvecs"):
            root = o.get("path", os.path.join(self.data_dir, "sift"))
            self.train_vectors = read_fvecs(os.path.join(root, o.get("base_file", "sift_base.fvecs")), o.get("base_limit"))
            self.test_vectors = read_fvecs(os.path.join(root, o.get("query_file", "sift_query.fvecs")), o.get("query_limit"))
            gt_path = os.path.join(root, o.get("groundtruth_file", "sift_groundtruth.ivecs"))
            if os.path.exists(gt_path) and not o.get("base_limit"):
                self.ground_truth = read_ivecs(gt_path, o.get("query_limit"))
                self.ground_truth_source = "file"
        else:
            raise ValueError(f"Unknown dataset '{self.name}' (available: random, sift1m_shape, glove50_shape, npy, fvecs)")
        if self.ground_truth is None:
            self._compute_ground_truth(gt_k, metric)

    # ------------------------------------------------------------------ ground truth
    def _compute_ground_truth(self, k: int, metric: str) -> None:
        train, test = self.train_vectors, self.test_vectors
        k = min(k, train.shape[0])
        mode = self.options.get("ground_truth", "auto")
        if mode == "cpu" or (mode == "auto" and float(train.shape[0]) * test.shape[0] <= _CPU_GT_LIMIT):
            gt = np.zeros((test.shape[0], k), dtype=np.int32)
            base = np.asarray(train, dtype=np.float32)
            if metric == "cosine":
                nb = np.linalg.norm(base, axis=1, keepdims=True)
                base = np.divide(base, nb, out=np.zeros_like(base), where=nb > 0)
            for i in range(test.shape[0]):
                q = np.asarray(test[i], dtype=np.float32)
                if metric == "l2":
                    gt[i] = np.argsort(np.linalg.norm(base - q[None, :], axis=1), kind="stable")[:k]
                else:
                    gt[i] = np.argsort(-(base @ q), kind="stable")[:k]
            self.ground_truth, self.ground_truth_source = gt, "numpy"
            return
        from ..indexes import GpuIndexFlat     # exact GPU search as ground-truth builder (dataset.py:858-964)
        index = GpuIndexFlat(train.shape[1], "l2" if metric == "l2" else "ip", normalize=metric == "cosine")
        index.add(train)
        _, idx = index.search(np.asarray(test, dtype=np.float32), k)
        self.ground_truth, self.ground_truth_source = idx.astype(np.int32), "gpu_exact"
