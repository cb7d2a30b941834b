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
if not os.path.exists(script):      # a bare script name: the copy vendored next to the oracle (GPU box) or the checkout
    for base in (os.path.join(ROOT, "oracle", "_ref"), os.environ.get("FFN_REFERENCE", "/root/reference")):
        if os.path.exists(os.path.join(base, sys.argv[1])):
            script = os.path.join(base, sys.argv[1])
            break
sys.argv = [script] + sys.argv[2:]
sys.path = [os.path.join(ROOT, "compat"), ROOT] + [p for p in sys.path
                                                     if os.path.abspath(p or ".") != os.path.dirname(script)]
if not os.environ.get("DISPLAY"):
    # headless box: train_image_regression.py / train_signal_regression.py preview with cv2.imshow when not on AzureML
    import cv2
    cv2.imshow = lambda *a, **k: None
    cv2.waitKey = lambda *a, **k: -1
if os.environ.get("FFN_REPORT_LAUNCHES"):
    # tests: prove that the script's work went through libffn_b200 (kernel launches issued by the library)
    import atexit

    def _report():
        from fourier_feature_nets_b200 import _lib
        print("FFN_LAUNCHES", _lib.launch_count() if _lib._lib is not None else 0)
    atexit.register(_report)
code = compile(open(script).read(), script, "exec")
globs = {"__name__": "__main__", "__file__": script}
exec(code, globs)
