"""Counterpart of the reference's sim/src/main.py:10-27:  python -m cuda_sph_b200.main [--mode BOX|PIPE] [-n N] ..."""
from __future__ import annotations

import argparse
import logging

from . import config
from .serializer import Saver
from .state_generator import StateGenerator


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default=config.SIM_MODE, choices=["BOX", "PIPE"])
    ap.add_argument("-n", "--particles", type=int, default=config.PARTICLE_COUNT)
    ap.add_argument("--duration", type=int, default=config.DURATION)
    ap.add_argument("--fps", type=int, default=config.FPS)
    ap.add_argument("--out", default=config.OUT_DIRNAME)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--steps-per-frame", type=int, default=1)
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO)
    if args.mode == "PIPE":
        params = config.pipe_params(args.particles, duration=args.duration, fps=args.fps)
        start = config.start_state_inside_pipe(args.particles, params.pipe, args.seed)
    else:
        params = config.box_params(args.particles, duration=args.duration, fps=args.fps)
        start = config.start_state_box_wall(args.particles, params.space_size, args.seed)
    saver = Saver(args.out, params, asynchronous=True)
    gen = StateGenerator(start, params, config.constants(args.mode), steps_per_frame=args.steps_per_frame)
    logging.info("Thread layout: grid size %d, block size %d", gen.sph_strategy.grid_size,
                 gen.sph_strategy.block_size)
    for state in gen:
        saver.save_next_state(state)
    saver.close()
    logging.info("Simulation finished.")


if __name__ == "__main__":
    main()
