#!/usr/bin/env python
"""Entry point with the reference's command line (scripts/run_full_benchmark.py:281-320):
    python scripts/run_full_benchmark.py --config configs/benchmark_config_smoke.yaml [--output-dir DIR]"""
import argparse
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from vectordb_retrieval_b200.harness import BenchmarkRunner  # noqa: E402


def main() -> int:
    parser = argparse.ArgumentParser(description="Run the vector retrieval benchmark on the B200 scan + top-k path")
    parser.add_argument("--config", type=str, default="configs/benchmark_config_smoke.yaml")
    parser.add_argument("--output-dir", type=str, default="benchmark_results")
    args = parser.parse_args()
    logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(name)s - %(levelname)s - %(message)s")
    results = BenchmarkRunner(args.config, args.output_dir).run()
    for ds, algs in results.items():
        for name, r in algs.items():
            print(f"{ds:>16s} {name:>12s} recall={r.get('recall', float('nan')):.4f} qps={r.get('qps', 0.0):.1f}")
    return 0 if results else 1


if __name__ == "__main__":
    sys.exit(main())
