"""Frame export in the reference's on-disk format (common/serializer/saver.py:17-46, loader.py:14-64), so the
reference's vis/ viewer and analize.py read our output unchanged:

    <root>/<out_dirname>/params.json                      sorted keys, indent 4
    <root>/<out_dirname>/{position,velocity,density}_<epoch>.npy     fp64, (N,3)/(N,3)/(N,), epoch counts from 0

`Saver(asynchronous=True)` writes frames on a background thread so the GPU loop does not wait for np.save.

Beyond the reference's files (all ignored by its Loader / viewer / analize.py, which open files by name):
    rng_states_<epoch>.npy        xoroshiro128+ states after that frame (PIPE-mode checkpoint frames) -> resume
    stats.jsonl                   one JSON line per frame: on-device reductions (analize.py:9-14 + neighbour histogram)
    <out_dirname>_preview/        the same run down-sampled to <= `preview_max_points` particles (every k-th id) with its
                                  own params.json, so the viewer (100 000-point cap, gl_point_field.py:11) opens it as is
"""
from __future__ import annotations

import dataclasses
import json
import os
import queue
import threading
from dataclasses import fields

import numpy as np

from .config import PARAMS_FILENAME
from .data_classes import Pipe, Segment, SimulationParameters, SimulationState


class Saver:
    def __init__(self, out_dirname: str, params: SimulationParameters, root: str | None = None,
                 asynchronous: bool = False, first_epoch: int = 0, preview_max_points: int = 0) -> None:
        self._dir = os.path.join(root or os.getcwd(), out_dirname)
        os.makedirs(self._dir, exist_ok=True)
        self._write_params(params, self._dir)
        self._epoch = int(first_epoch)
        self._stats_fh = None
        self._preview_dir, self._preview_stride = None, 1
        n = int(params.particle_count)
        if preview_max_points and n > preview_max_points:
            self._preview_stride = -(-n // int(preview_max_points))
            self._preview_dir = self._dir + "_preview"
            os.makedirs(self._preview_dir, exist_ok=True)
            small = dataclasses.replace(params, particle_count=-(-n // self._preview_stride))
            self._write_params(small, self._preview_dir)
        self._q: queue.Queue | None = None
        self._err: BaseException | None = None
        if asynchronous:
            self._q = queue.Queue(maxsize=4)
            self._thread = threading.Thread(target=self._drain, daemon=True)
            self._thread.start()

    @staticmethod
    def _write_params(params: SimulationParameters, directory: str) -> None:
        as_dict = dataclasses.asdict(params)
        for key, value in as_dict.items():
            if isinstance(value, np.ndarray):
                as_dict[key] = value.tolist()
        with open(os.path.join(directory, PARAMS_FILENAME), "w") as fh:
            json.dump(as_dict, fh, default=lambda o: o.__dict__, sort_keys=True, indent=4)

    def _write_frame(self, epoch: int, state: SimulationState, stats: dict | None = None) -> None:
        for name, value in vars(state).items():      # position / velocity / density (+ rng_states on checkpoint frames)
            np.save(os.path.join(self._dir, f"{name}_{epoch}"), value)
        if self._preview_dir:
            k = self._preview_stride
            for f in fields(SimulationState):
                np.save(os.path.join(self._preview_dir, f"{f.name}_{epoch}"), getattr(state, f.name)[::k])
        if stats is not None:
            if self._stats_fh is None:
                self._stats_fh = open(os.path.join(self._dir, "stats.jsonl"), "a")
            self._stats_fh.write(json.dumps(dict(stats, epoch=epoch)) + "\n")
            self._stats_fh.flush()

    def _drain(self) -> None:
        while True:
            item = self._q.get()
            if item is None:
                return
            try:
                self._write_frame(*item)
            except BaseException as exc:  # surfaced by the next save / close
                self._err = exc

    def save_next_state(self, state: SimulationState, stats: dict | None = None) -> None:
        if self._err:
            raise self._err
        if self._q is not None:
            self._q.put((self._epoch, state, stats))
        else:
            self._write_frame(self._epoch, state, stats)
        self._epoch += 1

    def close(self) -> None:
        if self._q is not None:
            self._q.put(None)
            self._thread.join()
            self._q = None
        if self._stats_fh is not None:
            self._stats_fh.close()
            self._stats_fh = None
        if self._err:
            raise self._err


class Loader:
    def __init__(self, out_dirname: str, root: str | None = None) -> None:
        self._dir = os.path.join(root or os.getcwd(), out_dirname)
        if not os.path.exists(self._dir):
            raise Exception(f"Directory ({self._dir}) does not exists! Could not load simulation!")
        with open(os.path.join(self._dir, PARAMS_FILENAME)) as fh:
            self._json = json.load(fh)

    def load_simulation_parameters(self) -> SimulationParameters:
        values = {}
        for f in fields(SimulationParameters):
            raw = self._json[f.name]
            if f.name == "pipe":
                values["pipe"] = Pipe([Segment(start_point=tuple(s["start_point"]), start_radius=s["start_radius"],
                                               end_radius=s["end_radius"], length=s["length"])
                                       for s in raw["segments"]])
            elif isinstance(raw, list):
                values[f.name] = np.asarray(raw)
            else:
                values[f.name] = raw
        return SimulationParameters(**values)

    def load_rng_states(self, epoch: int):
        """xoroshiro128+ states saved with a checkpoint frame, or None (no reference counterpart: it cannot resume)."""
        path = os.path.join(self._dir, f"rng_states_{epoch}.npy")
        return np.load(path) if os.path.exists(path) else None

    def load_simulation_state(self, epoch: int) -> SimulationState:
        return SimulationState(**{f.name: np.load(os.path.join(self._dir, f"{f.name}_{epoch}.npy"))
                                  for f in fields(SimulationState)})
