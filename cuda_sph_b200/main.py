"""Counterpart of the reference's sim/src/main.py:10-27:  python -m cuda_sph_b200.main [--mode BOX|PIPE] [-n N] ...

Beyond the reference's loop: --steps-per-frame (sub-stepping), --device-start (seeded start state
generated on the GPU), --resume K (continue from frame K of an earlier run in the same output directory, PIPE mode from
a checkpoint frame that carries RNG states), --preview N (down-sampled copy for the viewer), --stats (per-frame on-device
reductions to stats.jsonl)."""
from __future__ import annotations

import argparse
import logging

from . import config
from .serializer import Loader, Saver
from .state_generator import StateGenerator


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default=config.SIM_MODE, choices=["BOX", "PIPE"])
    ap.add_argument("-n", "--particles", type=int, default=config.PARTICLE_COUNT)
    ap.add_argument("--duration", type=int, default=config.DURATION)
    ap.add_argument("--fps", type=int, default=config.FPS)
    ap.add_argument("--out", default=config.OUT_DIRNAME)
    ap.add_argument("--root", default=None, help="directory that holds the output directory (default: cwd)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--steps-per-frame", type=int, default=1)
    ap.add_argument("--device-start", action="store_true", help="generate the start state on the GPU (hashed draws)")
    ap.add_argument("--resume", type=int, default=None, metavar="K", help="continue after frame K of the run in --out")
    ap.add_argument("--checkpoint-every", type=int, default=10, help="PIPE mode: frames between RNG checkpoints")
    ap.add_argument("--preview", type=int, default=100000, metavar="N",
                    help="also write <out>_preview with at most N particles per frame (0 = off)")
    ap.add_argument("--stats", action="store_true", help="write per-frame on-device reductions to stats.jsonl")
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO)
    if args.mode == "PIPE":
        params = config.pipe_params(args.particles, duration=args.duration, fps=args.fps)
    else:
        params = config.box_params(args.particles, duration=args.duration, fps=args.fps)
    first, rng, start, generate = 0, None, None, None
    if args.resume is not None:
        loader = Loader(args.out, root=args.root)
        start = loader.load_simulation_state(args.resume)
        rng = loader.load_rng_states(args.resume)
        if args.mode == "PIPE" and rng is None:
            raise SystemExit(f"frame {args.resume} carries no RNG states: resume from a checkpoint frame "
                             f"(every {args.checkpoint_every} frames)")
        first = args.resume + 1
    elif args.device_start:
        generate = ("pipe" if args.mode == "PIPE" else "box_wall", args.seed)
    elif args.mode == "PIPE":
        start = config.start_state_inside_pipe(args.particles, params.pipe, args.seed)
    else:
        start = config.start_state_box_wall(args.particles, params.space_size, args.seed)
    saver = Saver(args.out, params, root=args.root, asynchronous=True, first_epoch=first,
                  preview_max_points=args.preview)
    gen = StateGenerator(start, params, config.constants(args.mode), steps_per_frame=args.steps_per_frame,
                         first_frame=first, rng_states=rng, generate=generate, checkpoint_every=args.checkpoint_every)
    logging.info("Thread layout: grid size %d, block size %d", gen.sph_strategy.grid_size,
                 gen.sph_strategy.block_size)
    for state in gen:
        saver.save_next_state(state, gen.current_stats if args.stats else None)
    saver.close()
    logging.info("Simulation finished.")


if __name__ == "__main__":
    main()
