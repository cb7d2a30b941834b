"""Run one of the reference's scripts UNCHANGED against the B200 build:

    python tools/run_reference_script.py /path/to/reference/orbit_video.py model.pt 400 out --num-frames 8

``compat/`` (import name ``fourier_feature_nets`` + a ``scenepic`` stand-in) is put first on ``sys.path``;
the script's own directory, which ``runpy`` would otherwise put first and which holds the reference package,
is removed."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
script = os.path.abspath(sys.argv[1])
sys.argv = [script] + sys.argv[2:]
sys.path = [os.path.join(ROOT, "compat"), ROOT] + [p for p in sys.path
                                                     if os.path.abspath(p or ".") != os.path.dirname(script)]
if not os.environ.get("DISPLAY"):
    # headless box: train_image_regression.py / train_signal_regression.py preview with cv2.imshow when not on AzureML
    import cv2
    cv2.imshow = lambda *a, **k: None
    cv2.waitKey = lambda *a, **k: -1
code = compile(open(script).read(), script, "exec")
globs = {"__name__": "__main__", "__file__": script}
exec(code, globs)
