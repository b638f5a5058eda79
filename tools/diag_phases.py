# prints the level-set loop's diagnostics (running lanes per iteration, lanes per phase) for c2 and c4: VDBRT_DEBUG_TILES=1 python tools/diag_phases.py
import sys
sys.path.insert(0, '/root/repo')
from openvdb_b200 import api
ctx = api.Context(0)
g = ctx.build_torus(650.0, 325.0)
print('c2', ctx.count_levelset(g, api.vdb_render_camera(1920, 1080, (0, 1.5 * 650, 3 * (650 + 325.0)), (0, 0, 0))).as_dict(), flush=True)
g.free()
g = ctx.build_spheres(api.random_spheres(10000, 20240607, 1988.0, 10.0, 60.0))
print('c4', ctx.count_levelset(g, api.vdb_render_camera(3840, 2160, (0.0, 0.0, 3 * 2048.0), (0.0, 0.0, 0.0))).as_dict(), flush=True)
