"""Copy the reference's own test clip (BASELINE.json configs[0]: ``assets/aicamera_test_clip.mp4``,
960x540, 500 frames, H.264) into tests/golden/ so that the C1 parity test and ``bench.py --config clip``
can read it on the GPU box, where /root/reference does not exist.  The clip is INPUT DATA (frames), not
source code; it is decoded with cv2.VideoCapture exactly as the reference does
(/root/reference/src/aicamera_tracker.py:118,170).

    python tests/golden/make_clip_fixture.py [/root/reference]
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    src = os.path.join(ref, "assets", "aicamera_test_clip.mp4")
    dst = os.path.join(HERE, "aicamera_test_clip.mp4")
    shutil.copyfile(src, dst)
    os.chmod(dst, 0o644)
    import cv2
    cap = cv2.VideoCapture(dst)
    n, h = 0, hashlib.sha256()
    first = None
    while True:
        ok, f = cap.read()
        if not ok:
            break
        if n < 8:
            h.update(f.tobytes())  # decoded pixels of the first frames: pins the decoder as well as the file
        if first is None:
            first = f.shape
        n += 1
    meta = {"file": "aicamera_test_clip.mp4", "sha256": hashlib.sha256(open(dst, "rb").read()).hexdigest(),
            "frames": n, "frame_shape": list(first), "fps": cap.get(cv2.CAP_PROP_FPS),
            "decoded_sha256_first8": h.hexdigest(), "opencv": cv2.__version__,
            "source": "abdur75648/AI-Camera assets/aicamera_test_clip.mp4"}
    with open(os.path.join(HERE, "clip_meta.json"), "w") as fo:
        json.dump(meta, fo, indent=1)
    print(meta)


if __name__ == "__main__":
    main()
