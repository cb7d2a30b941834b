"""TEST INFRASTRUCTURE -- run one of the reference's scripts UNCHANGED against the reference's OWN package
(``oracle/_ref`` or ``/root/reference``), behind the module stand-ins of ``oracle/reference.py``:

    python oracle/run_ref_script.py orbit_video.py model.pt 48 out --device cpu --num-frames 3

The counterpart of ``tools/run_reference_script.py`` (same scripts against the B200 build): the pair gives the
frame-level parity tests their two arms."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import reference  # noqa: E402

loc = reference.location()
if not loc:
    raise SystemExit("reference not available (oracle/_ref is made by __graft_entry__.build())")
reference.install_shims()
script = sys.argv[1] if os.path.isabs(sys.argv[1]) else os.path.join(loc, sys.argv[1])
sys.argv = [script] + sys.argv[2:]
sys.path = [loc] + [p for p in sys.path if os.path.abspath(p or ".") != os.path.dirname(HERE)]
if not os.environ.get("DISPLAY"):
    import cv2
    cv2.imshow = lambda *a, **k: None
    cv2.waitKey = lambda *a, **k: -1
code = compile(open(script).read(), script, "exec")
exec(code, {"__name__": "__main__", "__file__": script})
