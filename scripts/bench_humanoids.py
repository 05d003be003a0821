#!/usr/bin/env python3
"""Timing of the humanoid configurations (BASELINE configs 3 / 4 without and with the barrier):
one JSON line per configuration (median of 5 regions of 5 graph-replayed launches)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench


def main():
    dev = torch.device("cuda", 0)
    peak, _ = bench.load_peaks()
    for c in bench.humanoid_configs(torch, dev, peak):
        print(json.dumps(c), flush=True)


if __name__ == "__main__":
    main()
