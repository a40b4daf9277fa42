"""Golden fixture for the pose I/O row (f-4): the head of the reference's own initial trajectory
(/root/reference/data/kitti_init_poor/00.txt, the file config/kitti_stereo.cfg:6 points apps/run_kitti at) and its
line count.  DATA of the reference, not source; generated in the build container, where the reference tree exists.

    python tests/golden/make_kitti_pose_golden.py
"""
import os

SRC = "/root/reference/data/kitti_init_poor/00.txt"
HERE = os.path.dirname(os.path.abspath(__file__))
lines = open(SRC).read().splitlines()
with open(os.path.join(HERE, "kitti_init_poor_00_head.txt"), "w") as f:
    f.write("\n".join(lines[:12]) + "\n")
with open(os.path.join(HERE, "kitti_init_poor_00_meta.txt"), "w") as f:
    f.write(f"{len([l for l in lines if l.strip()])}\n")
print("wrote", len(lines[:12]), "poses of", len(lines))

# the reference's own application configuration (config/kitti_stereo.cfg): the algorithm keys apps/run_kitti reads,
# kept verbatim so that the C++ class is exercised with the values the reference ships (patchRadius = 1, window of 5)
cfg = open("/root/reference/config/kitti_stereo.cfg").read()
with open(os.path.join(HERE, "kitti_stereo.cfg"), "w") as f:
    f.write(cfg)
print("wrote kitti_stereo.cfg,", len(cfg.splitlines()), "lines")
