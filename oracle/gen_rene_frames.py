"""Build-container script: camera / light poses of the reference's OWN ReNe fixture
(/root/reference/dataset_rene/savannah/train_transforms.json, 1628 frames = 44 cameras x 37 lights) pushed through the
reference's own pose conventions (projects/neuralangelo/data.py:121-151 get_camera / preprocess_camera / _gl_to_cv,
projects/NeuralLumen/data.py:30-44 get_light, projects/nerf/utils/camera.py Pose.invert), stored as
tests/golden/rene_savannah_frames.npz so that bench.py's rene_savannah_b workload (BASELINE.json configs[3], SURVEY.md
section 8d "C3") can draw its rays from the real frames on the GPU box, where /root/reference does not exist.

    python oracle/gen_rene_frames.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402


def main():
    ref = ref_import.load()
    camera = ref.camera
    meta = json.load(open("/root/reference/dataset_rene/savannah/train_transforms.json"))
    H, W = 270, 360  # rene_savannah_b.yaml:38-40
    # neuralangelo/data.py:39-42: `hasattr` on a dict is always False -> centre 0 / radius 1 whatever the JSON says
    center, scale = np.zeros(3), 1.0
    gl2cv = torch.tensor([1, -1, -1, 1], dtype=torch.float32)

    def w2c(mat):
        c2w = torch.tensor(mat, dtype=torch.float32) * gl2cv
        c2w[:3, -1] -= torch.from_numpy(center).float()
        c2w[:3, -1] /= scale
        return camera.Pose().invert(c2w[:3])

    intr = torch.tensor([[meta["fl_x"], meta["sk_x"], meta["cx"]], [meta["sk_y"], meta["fl_y"], meta["cy"]],
                         [0, 0, 1]]).float()
    intr[0] *= W / meta["w"]
    intr[1] *= H / meta["h"]
    frames = meta["frames"]
    pose = torch.stack([w2c(f["transform_matrix"]) for f in frames])
    pose_light = torch.stack([w2c(f["transform_matrix_light"]) for f in frames])
    out = os.path.join(ROOT, "tests", "golden", "rene_savannah_frames.npz")
    np.savez_compressed(out, pose=pose.numpy(), pose_light=pose_light.numpy(), intr=intr.numpy(),
                        camera_index=np.array([f["camera_index"] for f in frames], np.int16),
                        light_index=np.array([f["light_index"] for f in frames], np.int16),
                        image_size=np.array([H, W], np.int32),
                        bounding_box_aabb=np.array(meta["bounding_box_aabb"], np.float32))
    print(out, pose.shape, os.path.getsize(out))


if __name__ == "__main__":
    main()
