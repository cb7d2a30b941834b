"""TEST / BENCH INFRASTRUCTURE ONLY -- the REAL reference, vendored at build time.

``vendor()`` (called by ``__graft_entry__.build()`` in the build container, where ``/root/reference`` is
mounted) copies the reference's package and the scripts of the BASELINE configs, unmodified, into
``oracle/_ref/`` -- git-ignored (never in history), not gpurun-ignored (it travels to the GPU box like the built
``.so``).  ``import_reference()`` imports that copy (or ``/root/reference`` itself) as ``fourier_feature_nets``
behind stand-ins for the four packages that are absent from this image and do no arithmetic on the hot path
(SURVEY.md section 8c): ``scenepic`` (here WITH the camera maths ``utils.orbit`` needs, restated from scenepic's
documented behaviour), ``matplotlib.pyplot``, ``progress.bar``, ``trimesh``.

Users: ``bench.py --impl reference`` / its ``torch_gpu_baseline`` leg (the reference's own ``Raycaster`` timed on
the host cores / the same GPU), ``tests/`` (fixtures, script runs).  The product never imports this module.
"""
import os
import shutil
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("FFN_REFERENCE", "/root/reference")
REF_DIR = os.path.join(HERE, "_ref")

SCRIPTS = ["train_nerf.py", "orbit_video.py", "train_tiny_nerf.py", "train_voxels.py", "train_image_regression.py"]
EXTRA = [os.path.join("data", "cat.jpg"), os.path.join("docs", "ray_data.tsv"), "LICENSE"]


def vendor(force: bool = False) -> str:
    """Copy the reference package + config scripts into oracle/_ref (no-op where /root/reference is absent)."""
    if not os.path.isdir(os.path.join(REF_SRC, "fourier_feature_nets")):
        return REF_DIR if os.path.isdir(REF_DIR) else ""
    stamp = os.path.join(REF_DIR, ".vendored")
    if os.path.exists(stamp) and not force:
        return REF_DIR
    if os.path.isdir(REF_DIR):
        for root, dirs, files in os.walk(REF_DIR):      # an earlier copy may carry the mount's read-only modes
            os.chmod(root, 0o755)
        shutil.rmtree(REF_DIR)
    os.makedirs(REF_DIR)
    pkg = sorted(f for f in os.listdir(os.path.join(REF_SRC, "fourier_feature_nets")) if f.endswith(".py"))
    for name in [os.path.join("fourier_feature_nets", f) for f in pkg] + SCRIPTS + EXTRA:
        dst = os.path.join(REF_DIR, name)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF_SRC, name), dst)       # contents only: no read-only mode bits
    with open(stamp, "w") as f:
        f.write("copied unmodified from %s by oracle/reference.py; not tracked\n" % REF_SRC)
    return REF_DIR


def location() -> str:
    """Directory holding the reference (``fourier_feature_nets/`` + scripts): oracle/_ref, else /root/reference."""
    for d in (REF_DIR, REF_SRC):
        if os.path.isdir(os.path.join(d, "fourier_feature_nets")):
            return d
    return ""


def available() -> bool:
    return bool(location())


# ---------------------------------------------------------------------------------------------------------
# stand-ins for the absent third-party modules
# ---------------------------------------------------------------------------------------------------------
class _Any:
    """Accepts any construction / attribute / call (visualisation objects nobody reads)."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Any()

    def __call__(self, *a, **k):
        return _Any()

    def __iter__(self):
        return iter(())

    def save_as_html(self, path, *a, **k):
        with open(path, "w") as f:
            f.write("<html><body>scenepic is not installed: no visualisation</body></html>")


class Transforms:
    """The scenepic transforms the reference calls outside visualisation code: 4x4 float32, column vectors."""

    @staticmethod
    def scale(s):
        m = np.eye(4, dtype=np.float32)
        m[:3, :3] *= np.asarray(s, np.float32)
        return m

    @staticmethod
    def translate(v):
        m = np.eye(4, dtype=np.float32)
        m[:3, 3] = np.asarray(v, np.float32)
        return m

    @staticmethod
    def rotation_about_x(angle):
        c, s = np.float32(np.cos(angle)), np.float32(np.sin(angle))
        m = np.eye(4, dtype=np.float32)
        m[1, 1], m[1, 2], m[2, 1], m[2, 2] = c, -s, s, c
        return m

    @staticmethod
    def rotation_matrix_from_axis_angle(axis, angle):
        """Rodrigues' formula about the normalised axis."""
        a = np.asarray(axis, np.float32).reshape(3)
        a = a / np.linalg.norm(a)
        x, y, z = a
        c, s = np.float32(np.cos(angle)), np.float32(np.sin(angle))
        t = np.float32(1) - c
        m = np.eye(4, dtype=np.float32)
        m[:3, :3] = np.array([[t * x * x + c, t * x * y - s * z, t * x * z + s * y],
                              [t * x * y + s * z, t * y * y + c, t * y * z - s * x],
                              [t * x * z - s * y, t * y * z + s * x, t * z * z + c]], np.float32)
        return m

    @staticmethod
    def look_at_rotation(center, look_at, up_dir):
        """OpenGL look-at: rows = camera x (right), y (up), z (backward) axes in world coordinates."""
        center, look_at, up_dir = [np.asarray(v, np.float32).reshape(3) for v in (center, look_at, up_dir)]
        z = center - look_at
        z = z / np.linalg.norm(z)
        x = np.cross(up_dir, z)
        x = x / np.linalg.norm(x)
        y = np.cross(z, x)
        m = np.eye(4, dtype=np.float32)
        m[0, :3], m[1, :3], m[2, :3] = x, y, z
        return m

    def __getattr__(self, name):        # gl_projection etc.: visualisation only
        return _Any()


class Camera:
    """``sp.Camera(center, look_at=(0,0,0), up_dir=(0,1,0), ...)``: OpenGL convention (looks down -z, +y up);
    ``world_to_camera = look_at_rotation @ translate(-center)``, ``camera_to_world`` its inverse."""

    def __init__(self, center=(0, 0, 4), look_at=(0, 0, 0), up_dir=(0, 1, 0), *args, **kwargs):
        center = np.asarray(center, np.float32)
        if center.shape == (4, 4):      # Camera(world_to_camera, projection): visualisation only
            self.world_to_camera = center
        else:
            self.world_to_camera = Transforms.look_at_rotation(center, look_at, up_dir) @ Transforms.translate(-center)
        self.camera_to_world = np.linalg.inv(self.world_to_camera).astype(np.float32)


class Bar:
    """progress.bar.Bar: prints nothing."""

    def __init__(self, *a, **k):
        self.suffix = ""
        self.index = 0
        self.max = k.get("max", 100)

    def next(self, *a, **k):
        self.index += 1

    def finish(self):
        pass

    def writeln(self, line):
        pass

    def update(self):
        pass

    @property
    def elapsed(self):
        return 0

    @property
    def eta(self):
        return 0


def install_shims():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    if "scenepic" not in sys.modules or not hasattr(sys.modules["scenepic"], "__ffn_shim__"):
        sp = mod("scenepic", Camera=Camera, Transforms=Transforms, Scene=_Any, Mesh=_Any, Colors=_Any(), Shading=_Any,
                 __ffn_shim__=True)

        def _sp_getattr(name):
            if name.startswith("__"):
                raise AttributeError(name)
            return _Any()
        sp.__getattr__ = _sp_getattr
    try:
        import matplotlib.pyplot  # noqa: F401
    except ImportError:
        mpl = mod("matplotlib")
        mpl.pyplot = mod("matplotlib.pyplot", get_cmap=lambda *a, **k: (lambda x: np.zeros(np.shape(x) + (4,))),
                         Axes=_Any, Figure=_Any)
    try:
        import progress.bar  # noqa: F401
    except ImportError:
        prog = mod("progress")
        prog.bar = mod("progress.bar", Bar=Bar, ChargingBar=Bar)
    try:
        import trimesh  # noqa: F401
    except ImportError:
        mod("trimesh")


def import_reference():
    """``import fourier_feature_nets`` = the unmodified reference package (from oracle/_ref or /root/reference)."""
    loc = location()
    if not loc:
        raise ImportError("the reference is neither vendored (oracle/_ref, made by __graft_entry__.build() in the "
                          "build container) nor mounted (%s)" % REF_SRC)
    install_shims()
    cur = sys.modules.get("fourier_feature_nets")
    if cur is not None and os.path.realpath(getattr(cur, "__file__", "")).startswith(os.path.realpath(loc)):
        return cur
    if cur is not None:
        raise ImportError("another 'fourier_feature_nets' (%s) is already imported in this process" % cur.__file__)
    sys.path.insert(0, loc)
    try:
        import fourier_feature_nets as ref
    finally:
        sys.path.remove(loc)
    assert os.path.realpath(ref.__file__).startswith(os.path.realpath(loc)), ref.__file__
    return ref


if __name__ == "__main__":
    print(vendor(force="--force" in sys.argv) or "reference not available")
