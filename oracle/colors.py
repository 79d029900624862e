"""oracle/colors.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement (fp32, operation for operation) of the speed-gradient colours the reference's CPU backend
uploads next to the positions:

* ``speed_normalized``: FluidSimulation::getSpeedNormalzied, engine/physics/physicsWorld.cc:178-182 --
  ``glm::clamp(glm::length(velocity[i]), 0.0f, 1.5f) / 1.5f`` with glm::length = sqrt(dot) and glm's vec3 dot
  ``(x*x + y*y) + z*z`` (exts/glm/glm/detail/func_geometric.inl:48-55);
* ``speed_colors``: FluidSimCPU::updateColors, projects/Simulation/code/simulations/fluidSimCPU.cc:100-125, with
  the gradient stops of fluidSimCPU.h:26-29.

The reference's backend needs a GL context and cannot be compiled here, so this restatement is pinned by
hand-checked known answers (tests/test_oracle.py) rather than by the reference's own output.
"""
import numpy as np

F = np.float32
COLOR1 = np.array([0.0, 0.75, 1.0, 1.0], F)     # fluidSimCPU.h:26-29
COLOR2 = np.array([0.0, 1.0, 0.0, 1.0], F)
COLOR3 = np.array([1.0, 1.0, 0.0, 1.0], F)
COLOR4 = np.array([1.0, 0.0, 0.0, 1.0], F)
B1, B2 = F(0.33), F(0.66)                        # fluidSimCPU.cc:110-111


def speed_normalized(vel):
    v = np.asarray(vel, F)
    dot = (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2]        # all fp32
    length = np.sqrt(dot, dtype=F)
    return (np.minimum(np.maximum(length, F(0.0)), F(1.5)) / F(1.5)).astype(F)


def speed_colors(vel):
    n = speed_normalized(vel)
    out = np.empty((n.size, 4), F)
    for sel, a, lo, hi in (
            (n <= B1, n / B1, COLOR1, COLOR2),                                           # :113-115
            ((n > B1) & (n <= B2), (n - B1) / (B2 - B1), COLOR2, COLOR3),                # :116-119
            (n > B2, (n - B2) / (F(1.0) - B2), COLOR3, COLOR4)):                         # :120-123
        a = a.astype(F)[sel]
        out[sel] = ((F(1.0) - a)[:, None] * lo[None, :] + a[:, None] * hi[None, :]).astype(F)
    return out
