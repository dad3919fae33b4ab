"""A/B timing of kernel variants (development): each variant runs tools/ncu_target.py in a subprocess with
different RT_B200_* switches and prints Mpaths/s for a few scenes."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENES = [(8, 800, 64, 0), (6, 600, 64, 50), (0, 400, 256, 50), (3, 800, 64, 0), (7, 600, 64, 0)]
VARIANTS = [{}, {"RT_B200_KERNEL": "1"}, {"RT_B200_MIN_BLOCKS": "5"}, {"RT_B200_SHADE_MIN": "16"}, {"RT_B200_NO_BOX": "1"},
            {"RT_B200_NO_HOIST": "1"}]
if len(sys.argv) > 1:
    import json
    VARIANTS = json.loads(sys.argv[1])
if len(sys.argv) > 2:
    keep = [int(x) for x in sys.argv[2].split(",")]
    SCENES = [s for s in SCENES if s[0] in keep]
if len(sys.argv) > 3:      # spp override
    SCENES = [(a, b, int(sys.argv[3]), d) for a, b, _, d in SCENES]
for var in VARIANTS:
    for scene, width, spp, depth in SCENES:
        env = dict(os.environ, **var)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_target.py"), "--scene", str(scene), "--width", str(width),
                              "--spp", str(spp), "--reps", "3", "--depth", str(depth)], env=env, capture_output=True, text=True)
        last = [l for l in out.stdout.splitlines() if l.startswith("rep")]
        print(var, "scene", scene, last[-1] if last else out.stderr[-300:], flush=True)
