"""Oracle (test infrastructure): NumPy twin of the synthetic vector environment
(madrona-learn_b200/csrc/env.cu; SURVEY 8d "Synthetic inputs").  Bit-exact by construction:
noise is threefry bits -> exact int-to-float conversion, arithmetic is unfused float32."""
import numpy as np

from .prng import threefry2x32

F32 = np.float32
U32 = np.uint32


def _noise(bits):
    u = (bits >> U32(8)).astype(F32) * F32(5.9604644775390625e-08)
    return (u * F32(3.4641016151377544) + F32(-1.7320508075688772)).astype(F32)


class SyntheticEnv:
    def __init__(self, N, D=64, A=6, seed=0, p_done=1.0 / 64):
        self.N, self.D, self.A, self.seed, self.p_done = N, D, A, U32(seed & 0xFFFFFFFF), F32(p_done)
        self.t = 0
        e = np.arange(N * D, dtype=np.uint64)
        x0, _ = threefry2x32(self.seed, (e >> np.uint64(32)).astype(U32), e.astype(U32),
                             np.full(N * D, 0xFFFFFFFF, U32))
        self.obs = _noise(x0).reshape(N, D)

    def step(self, actions):
        N, D, t = self.N, self.D, self.t
        n = np.arange(N, dtype=np.uint64)
        if self.p_done < 0:
            done = ((t + n) % 61) == 0
        else:
            x0, _ = threefry2x32(self.seed, U32(0x9E3779B9), n.astype(U32),
                                 np.full(N, 0x80000000 | t, U32))
            done = ((x0 >> U32(8)).astype(F32) * F32(5.9604644775390625e-08)) < self.p_done
        e = np.arange(N * D, dtype=np.uint64)
        x0, _ = threefry2x32(self.seed, (e >> np.uint64(32)).astype(U32), e.astype(U32),
                             np.full(N * D, t, U32))
        xi = _noise(x0).reshape(N, D)
        o = self.obs
        rewards = ((o[:, 0] * (actions[:, 0].astype(F32) + F32(-1.5))) * F32(0.1)).astype(F32)
        nxt = (F32(0.9) * o + F32(0.1) * xi).astype(F32)
        self.obs = np.where(done[:, None], xi, nxt).astype(F32)
        self.t += 1
        return self.obs, rewards, done
