#!/usr/bin/env python
"""Regenerate tests/golden/*.npz.

The reference is Julia and cannot run in this image, so these vectors are NOT reference outputs: they are outputs of
the CPU oracle (oracle/sde_oracle.c) at the commit that introduced them, frozen so that (i) a later change of the
oracle is detected, and (ii) the CUDA path can be checked without building the oracle (the GPU tests compare
bit for bit in STRICT math).  Inputs are derived from the documented Philox normal stream (oracle/philox_spec.c), so the
files carry their own inputs."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import oracle as O

BS, HE = (0.01, 0.3), (0.01, 10.0, 0.3, 0.1, 0.4)


def main():
    out = {}
    # Wiener-driven trajectories (simulation.jl:119-134)
    for tag, model, p, x0, M, sid, nsteps in (("bs_em", "BlackScholes", BS, [100.0], 1, 0, 24), ("bs_mil", "BlackScholes", BS, [100.0], 1, 1, 24),
                                              ("heston_em", "Heston", HE, [100.0, 0.4], 2, 0, 32)):
        dt = 1.0 / nsteps
        z = O.normals_paths(2024, 1, 0, 6, nsteps, M)
        w = O.wiener_z(dt, z)
        out[tag + "_w"] = w
        out[tag + "_x"] = O.simulate_wiener(model, sid, p, 0.0, dt, x0, w)
        out[tag + "_terminal_from_z"] = O.sample_z(model, sid, p, 0.0, dt, x0, z)
    # coupled level pair (multilevel.jl:65-89): substeps 4, 5 coarse steps
    z = O.normals_paths(7, 3, 0, 5, 20, 2)
    xc, xf = O.multilevel_sample_z("Heston", 0, HE, 0.0, 0.1, O.level_dt(0.1, 4, 1), 4, [100.0, 0.4], z)
    out["mlmc_z"], out["mlmc_xc"], out["mlmc_xf"] = z, xc, xf
    # budget allocator (multilevel.jl:12-24)
    out["npaths_4_0_4_1000"] = O.npaths(4, 0, 4, 1000)
    out["npaths_2_0_8_2p34"] = O.npaths(2, 0, 8, 2 ** 34)
    # the normal stream itself
    out["normals_f64_seed5_stream1_path77"] = O.normals(5, 1, 77, 0, 64)
    out["normals_f32_seed5_stream1_path77"] = O.normals(5, 1, 77, 0, 64, "f32")
    np.savez(os.path.join(HERE, "oracle_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "oracle_vectors.npz"), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
