#!/usr/bin/env python
"""BASELINE.json configs[2]: "input_images/2400.jpg: 8 SAM-segmented objects".  GroundingDINO / SAM weights and the Gemini key
are not reachable from the build environment, so the segmentation stage cannot run.  This
script (build container only: it reads the reference checkout) cuts 8 deterministic crops out of the reference's input
image with FIXED boxes -- the 5 meshable instances (3 chairs, table, vase) + 3 half-object sub-crops -- and stores them as
small JPEGs under tests/golden/crops_2400/.  bench.py --crops 2400 loads them as RGBA with a rectangular (opaque) alpha,
i.e. what `ImageProcessorV2` would receive from stage 2, and runs configs[2]'s 8-objects-over-8-GPUs shape."""
import os

from PIL import Image

BOXES = [  # (name, x0, y0, x1, y1) in pixels of the 1500 x 1000 image
    ("chair_left", 160, 260, 690, 975), ("chair_back", 360, 95, 640, 530), ("chair_right", 885, 120, 1290, 710),
    ("table", 420, 220, 1080, 805), ("vase", 695, 45, 850, 320),
    ("table_top", 420, 220, 1080, 450), ("table_base", 600, 440, 890, 805), ("chair_left_seat", 300, 490, 690, 975),
]


def main():
    src = os.path.join(os.environ.get("R3G_REFERENCE", "/root/reference"), "input_images", "2400.jpg")
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "crops_2400")
    os.makedirs(out, exist_ok=True)
    im = Image.open(src).convert("RGB")
    for i, (name, x0, y0, x1, y1) in enumerate(BOXES):
        c = im.crop((x0, y0, x1, y1))
        c.thumbnail((384, 384), Image.Resampling.LANCZOS)
        path = os.path.join(out, f"{i}_{name}.jpg")
        c.save(path, quality=88)
        print(path, c.size, os.path.getsize(path))


if __name__ == "__main__":
    main()
