"""CPU: the synthetic SAPIEN-format scene writer produces the reference's on-disk layout (datasets/sapien.py:28-69,
datagen/data_utils.py:189-242): {split}/rgb/r_i.png RGBA + transforms.json {"focal", "frames": {"r_i": 4x4}}."""
import json
import os

import numpy as np

from oracle import ref_cpu as O


def test_synthetic_scene_layout(tmp_path, built_lib):
    from aon_b200 import data
    root = data.write_synthetic_scene(str(tmp_path / "scene"), (32, 24), n_train=3, n_val=1, n_test=2, seed=1)
    for split, n in (("train", 3), ("val", 1), ("test", 2)):
        meta = json.load(open(os.path.join(root, split, "transforms.json")))
        assert set(meta) == {"focal", "frames"} and len(meta["frames"]) == n
        assert abs(meta["focal"] - 0.5 * 24 / np.tan(np.radians(17.5))) < 1e-9
        files = sorted(os.listdir(os.path.join(root, split, "rgb")))
        assert files == sorted("r_%d.png" % i for i in range(n))
        for fn in files:
            c2w = np.array(meta["frames"][fn[:-4]])
            assert c2w.shape == (4, 4) and np.allclose(c2w[3], [0, 0, 0, 1])
            assert np.allclose(c2w[:3, :3].T @ c2w[:3, :3], np.eye(3), atol=1e-6)          # rotation
            assert abs(np.linalg.norm(c2w[:3, 3]) - 4.0) < 1e-6                             # data_gen.py:79 radius 4
            rgba = data._load_rgba(os.path.join(root, split, "rgb", fn), (32, 24))
            assert rgba.shape == (24 * 32, 4)
            a = rgba[:, 3]
            assert set(np.unique(a.numpy())) <= {0.0, 1.0} and 0.02 < a.mean() < 0.6        # object in view, background empty
            # the camera looks at the origin along -z of its own frame (ray_utils.py:86-88 convention)
            assert np.allclose(c2w[:3, 2], c2w[:3, 3] / 4.0, atol=1e-6)
