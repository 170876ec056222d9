import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from mujoco_ros2_simulation_b200 import engine, scenes
cs = scenes.clutter500()
sc = engine.Scene(cs.model, nworld=1, device=0)
sc.update_poses(cs.poses)
R = np.array([[0, 0, -1.0], [-1.0, 0, 0], [0, 1.0, 0]]).reshape(1, 9)
cp = np.array([[3.0, 0.0, 1.5]])
for env in ({}, {"B2_MESH_COOP_RAYS": "400000"}, {"B2_MESH_COOP_RAYS": "400000", "B2_MESH_COOP_G": "16"}, {"B2_MESH_COOP_RAYS": "400000", "B2_MESH_COOP_G": "8"}):
    for k in ("B2_MESH_COOP_RAYS", "B2_MESH_COOP_G"):
        os.environ.pop(k, None)
    os.environ.update(env)
    cam = engine.DepthCamera(scenes.depth_camera(), scene=sc, device=0)
    for _ in range(5):
        cam.render_to_cloud(cp, R, want=("depth", "cloud"))
    wall = []
    for _ in range(100):
        t0 = time.perf_counter()
        cam.render_to_cloud(cp, R, want=("depth", "cloud"))
        wall.append((time.perf_counter() - t0) * 1e6)
    print(env, "depth frame host p50 %.1f us" % np.median(wall))
    cam.close()
