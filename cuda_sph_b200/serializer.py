"""Frame export in the reference's on-disk format (common/serializer/saver.py:17-46, loader.py:14-64), so the
reference's vis/ viewer and analize.py read our output unchanged:

    <root>/<out_dirname>/params.json                      sorted keys, indent 4
    <root>/<out_dirname>/{position,velocity,density}_<epoch>.npy     fp64, (N,3)/(N,3)/(N,), epoch counts from 0

`Saver(asynchronous=True)` writes frames on a background thread so the GPU loop does not wait for np.save.
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
                 asynchronous: bool = False) -> None:
        self._dir = os.path.join(root or os.getcwd(), out_dirname)
        os.makedirs(self._dir, exist_ok=True)
        self._write_params(params)
        self._epoch = 0
        self._q: queue.Queue | None = None
        self._err: BaseException | None = None
        if asynchronous:
            self._q = queue.Queue(maxsize=4)
            self._thread = threading.Thread(target=self._drain, daemon=True)
            self._thread.start()

    def _write_params(self, params: SimulationParameters) -> None:
        as_dict = dataclasses.asdict(params)
        for key, value in as_dict.items():
            if isinstance(value, np.ndarray):
                as_dict[key] = value.tolist()
        with open(os.path.join(self._dir, PARAMS_FILENAME), "w") as fh:
            json.dump(as_dict, fh, default=lambda o: o.__dict__, sort_keys=True, indent=4)

    def _write_frame(self, epoch: int, state: SimulationState) -> None:
        for name, value in vars(state).items():
            np.save(os.path.join(self._dir, f"{name}_{epoch}"), value)

    def _drain(self) -> None:
        while True:
            item = self._q.get()
            if item is None:
                return
            try:
                self._write_frame(*item)
            except BaseException as exc:  # surfaced by the next save / close
                self._err = exc

    def save_next_state(self, state: SimulationState) -> None:
        if self._err:
            raise self._err
        if self._q is not None:
            self._q.put((self._epoch, state))
        else:
            self._write_frame(self._epoch, state)
        self._epoch += 1

    def close(self) -> None:
        if self._q is not None:
            self._q.put(None)
            self._thread.join()
            self._q = None
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

    def load_simulation_state(self, epoch: int) -> SimulationState:
        return SimulationState(**{f.name: np.load(os.path.join(self._dir, f"{f.name}_{epoch}.npy"))
                                  for f in fields(SimulationState)})
