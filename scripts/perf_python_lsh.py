#!/usr/bin/env python
"""The Python-LSH pipeline (LSHIndexer + LSHSearcher, reference src/algorithms/lsh.py) on the C3 shape: 1.2M x 50 cosine,
10k queries, k = 100.  Candidate generation on the device (vdb_lsh_candidates) against the host walk (NumPy restatement of
the reference's Counter loop, timed on a query sample), same results.  One JSON object per line."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vectordb_retrieval_b200.algorithms as A  # noqa: E402
from vectordb_retrieval_b200.harness.dataset import Dataset  # noqa: E402


def main():
    n, d, nq, k = 1_200_000, 50, 10_000, 100
    ds = Dataset("glove50_shape", options={"train_size": n, "test_size": nq, "ground_truth": "skip", "seed": 42})
    ds._clustered(d, n, nq, 64, 0.3)
    for tables, bits, mult in ((12, 18, 8.0), (12, 12, 8.0)):
        res = {}
        for mode, sample in (("device", nq), ("host", 400)):
            algo = A.get_algorithm_instance("Composite", d, name="lsh", metric="cosine",
                                            indexer=dict(type="LSHIndexer", num_tables=tables, hash_size=bits, seed=42),
                                            searcher=dict(type="LSHSearcher", candidate_multiplier=mult, fallback_to_bruteforce=True,
                                                          candidate_generation=mode))
            t = time.perf_counter()
            algo.build_index(ds.train_vectors)
            build_s = time.perf_counter() - t
            q = ds.test_vectors[:sample]
            algo.batch_search(q[:64], k)
            best = float("inf")
            for _ in range(2):
                t = time.perf_counter()
                D, I = algo.batch_search(q, k)
                best = min(best, time.perf_counter() - t)
            res[mode] = (D, I)
            print(json.dumps({"num_tables": tables, "hash_size": bits, "candidate_multiplier": mult, "candidate_generation": mode,
                              "queries": sample, "seconds": best, "qps": sample / best, "build_s": build_s}), flush=True)
        same = bool(np.array_equal(res["device"][1][:400], res["host"][1]) and np.array_equal(res["device"][0][:400], res["host"][0]))
        print(json.dumps({"num_tables": tables, "hash_size": bits, "device_equals_host_on_the_sample": same}), flush=True)


if __name__ == "__main__":
    main()
