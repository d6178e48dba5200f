"""``Measure`` — named per-iteration timers dumped as JSON (``examples/utils.py:120-192`` in the reference; defined
there but unused by the scripts).  Device-aware: ``stop`` can synchronise the current CUDA stream first."""
from __future__ import annotations

import json
import os
import shutil
import time

import torch


class Measure:
    def __init__(self, log_dir, sub_dir, sync_device=False):
        self.num_iters, self.self_iter = -1, -1
        self.begin = time.time(); self.total_time = -1; self.start_time = 0.0
        self.time_map, self.accuracy, self.num_samples = {}, 0.0, 0
        self.sync_device = sync_device and torch.cuda.is_available()
        self.log_path = os.path.join(log_dir, sub_dir)
        os.makedirs(log_dir, exist_ok=True)
        if os.path.exists(self.log_path):
            shutil.rmtree(self.log_path)
        os.mkdir(self.log_path)

    def set_begin_time(self, ts): self.begin = ts
    def get_begin_time(self): return self.begin

    def set_num_iters(self, n):
        assert n >= 0
        self.num_iters = n

    def next_iter(self):
        self.self_iter += 1
        self.time_map[self.self_iter] = {}

    def start(self, name):
        self.time_map[self.self_iter][name] = 0
        if self.sync_device:
            torch.cuda.synchronize()
        self.start_time = time.time()

    def stop(self, name):
        if self.time_map.get(self.self_iter, -1) != -1 and self.time_map[self.self_iter].get(name, -1) == 0:
            if self.sync_device:
                torch.cuda.synchronize()
            self.time_map[self.self_iter][name] = round(time.time() - self.start_time, 6)

    def set_accuracy(self, a): self.accuracy = a
    def add_samples(self, n): self.num_samples += n

    def reset(self, num_iters=-1):
        self.start_time, self.time_map, self.self_iter, self.accuracy, self.num_samples = 0.0, {}, -1, 0.0, 0
        if num_iters != -1:
            self.num_iters = num_iters

    def save_report(self):
        if self.num_iters == -1:
            print("[Error] Incorrect iteration number %d." % self.num_iters)
            return -1
        log = {"num_iters": self.num_iters, "time": self.time_map, "num_samples": self.num_samples,
               "total_time": time.time() - self.begin}
        if self.accuracy:
            log["accuracy"] = self.accuracy
        with open(os.path.join(self.log_path, "iter-%d.txt" % self.num_iters), "w") as fp:
            json.dump(log, fp)
        return 0
