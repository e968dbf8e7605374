"""ezrt_b200/data/bunny.npz: the Stanford bunny as an indexed mesh (unique vertices + faces in file order), recovered from
the bunny's 4968 encoded triangles of the committed P3 scene (tests/golden/p3_scene.npz = P3/main.cpp:690-701 run through
readObj/buildBVH).  It is an INPUT ASSET for the benchmark scene S-1M (201 bunny instances); the pose of P3's
transform is irrelevant because readObj normalises every mesh to its unit box again.  python tests/golden/make_bunny_asset.py"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def main():
    t = np.load(os.path.join(HERE, "p3_scene.npz"))["tris"].reshape(-1, 36)
    bunny = t[(t[:, 18:21].sum(1) == 0) & (t[:, 21] == 1.0)]  # white, non-emissive = the bunny (floor is grey, sphere emissive)
    assert bunny.shape[0] == 4968
    v = bunny[:, :9].reshape(-1, 3)
    verts, inv = np.unique(v, axis=0, return_inverse=True)
    faces = inv.reshape(-1, 3).astype(np.int32)
    np.savez_compressed(os.path.join(ROOT, "ezrt_b200", "data", "bunny.npz"), verts=verts.astype(np.float32), faces=faces)
    print("bunny.npz: %d vertices, %d faces" % (verts.shape[0], faces.shape[0]))


if __name__ == "__main__":
    main()
