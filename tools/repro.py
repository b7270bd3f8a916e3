import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from spear_b200 import particles, abi
import test_gpu_more as T
n = int(sys.argv[1])
ctx = particles.ParticleContext(initial_slot_capacity=1 << 16)
s = ctx.create_system((3, 1234, 5, 6))
e = s.add_effect(T.comp_big(n), abi.make_transform((1, 2, 3)))
dt = s.effect_info(e).timeStep
for f in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    s.update([e], abi.make_frame_args(dt, f * dt, f))
    ctx.synchronize()
    print(f, s.effect_info(e).aliveCount)
