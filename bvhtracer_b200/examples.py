"""Scene specifications of the reference's examples (examples/*.rs), as data.

The examples are the CALLERS of the hot path, not part of it: what matters here is that both arms (the GPU
engine and the CPU oracle in tests/) are fed exactly the same meshes, cameras and per-frame transforms.
Everything is plain numpy float32 arithmetic mirroring the f32 expressions of the Rust sources; matrices are
built downstream (host mirror / oracle) from (scale, translation, angle_x, angle_z).

Camera for every size other than 640x640: the example's own spec unchanged (fovy 90 deg, aspect 1.0), only
W and H change -- u = x / W, v = y / H (renderer.rs:358-361); see SURVEY.md 8(d).
"""
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

F = np.float32


@dataclass
class CameraSpec:
    """SymmetricFovSpec / BoxSpec + CameraAttitudeSpec (camera.rs:59-155, 730-766)."""
    position: Tuple[float, float, float]
    forward: Tuple[float, float, float]
    right: Tuple[float, float, float]
    up: Tuple[float, float, float]
    near: float
    fovy_deg: Optional[float] = 90.0          # SymmetricFovSpec when set
    aspect: float = 1.0
    box: Optional[Tuple[float, float, float, float]] = None   # (left, right, bottom, top) BoxSpec when set


@dataclass
class ObjectSpec:
    """One SceneObject: model index + Transform3::new(scale, translation, Rx(angle_x) * Rz(angle_z))."""
    model: int
    scale: Tuple[float, float, float] = (1.0, 1.0, 1.0)
    translation: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    angle_x: float = 0.0
    angle_z: float = 0.0
    with_transform: bool = True               # SceneObjectBuilder::with_transform called? (bounds stay empty otherwise)


@dataclass
class SceneSpec:
    name: str
    meshes: List[str]                         # asset names (assets/<name>.f32)
    camera: CameraSpec
    objects: List[ObjectSpec] = field(default_factory=list)
    default_size: Tuple[int, int] = (640, 640)     # SCREEN_WIDTH x SCREEN_HEIGHT of the example
    bench_size: Tuple[int, int] = (640, 640)       # size named by BASELINE.json for this config


def _normalize(v):
    v = np.asarray(v, F)
    m = np.sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2], dtype=F)
    return v / m


# ------------------------------------------------------------------------------------------ cube (C1)
def cube():
    """examples/cube.rs:25-80, frame 0 (before any update)."""
    pos = np.array([0, 4, 0], F)
    fwd = _normalize(np.zeros(3, F) - pos)
    cam = CameraSpec(position=tuple(pos), forward=tuple(fwd), right=(-1, 0, 0), up=(0, 0, 1), near=1.0)
    return SceneSpec("cube", ["cube.obj"], cam,
                     [ObjectSpec(0, scale=(2, 2, 2), translation=(-1, -1, -1))],
                     default_size=(640, 640), bench_size=(640, 640))


# ------------------------------------------------------------------------------------------ two_armadillos (C2)
def two_armadillos(frame="canonical"):
    """examples/two_armadillos.rs:25-105.

    frame="initial": both instances at identity (coincident geometry: a tie-breaking stress test).
    frame="canonical": rigid bodies at (-1.3, 0, 0) / (+1.3, 0, 0) with identity orientation, i.e. what
    update_transform (scene_object.rs:52-58) yields after the first physics step up to a negligible rotation
    (SURVEY.md 8d C2).  Transforms are explicit, so no physics restatement is needed.
    """
    cam = CameraSpec(position=(0, 1, -2.5), forward=(0, 0, 1), right=(1, 0, 0), up=(0, 1, 0), near=2.0)
    if frame == "initial":
        objs = [ObjectSpec(0), ObjectSpec(0)]
    else:
        objs = [ObjectSpec(0, translation=(-1.3, 0, 0)), ObjectSpec(0, translation=(1.3, 0, 0))]
    return SceneSpec("two_armadillos", ["armadillo.tri"], cam, objs, bench_size=(1920, 1080))


# ------------------------------------------------------------------------------------------ sixteen_* (C3, C4)
class GridAnimation:
    """The closed-form `Physics` of examples/sixteen_armadillos.rs:21-44, 75-122, 132-163 (identical in
    trippy_teapots.rs).  frame 0 = the scene as constructed; frame k >= 1 = after k calls of update(1/60)."""

    HEIGHT_INIT = [5, 4, 3, 2, 1, 5, 4, 3, 5, 4, 3, 2, 1, 5, 4, 3]

    def __init__(self, elapsed=1.0 / 60.0):
        self.elapsed = F(elapsed)                       # `elapsed as f32`
        self.angle = np.zeros(16, F)
        self.position_init = np.zeros((16, 3), F)
        self.height = np.zeros(16, F)
        self.speed = np.zeros(16, F)
        self.angular_velocity = np.zeros(16, F)
        self.acceleration = np.zeros(16, F)
        self.direction = np.full(16, -1, F)
        i = 0
        for x in range(4):
            for y in range(4):
                even = ((x + y) & 1) == 0
                # `((i * 13) & 7 + 2) as f32 * 0.10` parses as (i*13) & (7+2)
                self.angular_velocity[i] = F((i * 13) & 9) * F(0.10) if even else F(0)
                horizontal = np.array([(F(x) - F(1.5)) * F(2.5), 0, (F(y) - F(1.5)) * F(2.5)], F)
                vertical = np.zeros(3, F) if even else np.array([0, self.HEIGHT_INIT[i // 2], 0], F)
                self.position_init[i] = horizontal + vertical
                self.height[i] = vertical[1]
                self.acceleration[i] = F(0) if even else F(9.8)
                i += 1
        self.frame = 0
        # transforms the scene currently holds (frame 0: as constructed, rotation angle 0)
        self.current = [(tuple(self.position_init[i]), F(0)) for i in range(16)]

    def update(self):
        # sixteen_armadillos.rs:133-144: transforms from the CURRENT physics state ...
        self.current = []
        for i in range(16):
            tr = self.position_init[i] + np.array([0, self.height[i], 0], F)
            self.current.append((tuple(tr), F(self.angle[i])))
        # ... then integrate (:146-160)
        e = self.elapsed
        for i in range(16):
            self.angle[i] = self.angle[i] + self.angular_velocity[i] * e
            self.speed[i] = self.speed[i] + self.acceleration[i] * e
            self.height[i] = self.height[i] + self.direction[i] * self.speed[i] * e
            if self.height[i] < F(-3):
                self.height[i] = F(-3) + F(0.01)
                self.direction[i] = -self.direction[i]
                self.speed[i] = F(0.2)
            elif self.height[i] > self.position_init[i][1]:
                self.height[i] = self.position_init[i][1] - F(0.01)
                self.direction[i] = -self.direction[i]
        self.frame += 1

    def objects(self):
        return [ObjectSpec(0, scale=(0.75, 0.75, 0.75), translation=tr, angle_x=float(a), angle_z=float(a))
                for tr, a in self.current]


def sixteen_armadillos(frame=0):
    """examples/sixteen_armadillos.rs (C3): 16 instances of one armadillo BLAS; TLAS rebuilt per frame."""
    cam = CameraSpec(position=(0, 1, -5.5), forward=(0, 0, 1), right=(1, 0, 0), up=(0, 1, 0), near=2.0)
    anim = GridAnimation()
    for _ in range(frame):
        anim.update()
    return SceneSpec("sixteen_armadillos", ["armadillo.tri"], cam, anim.objects(), bench_size=(3840, 2160))


def trippy_teapots(frame=0):
    """examples/trippy_teapots.rs (C4): same grid/animation, teapot BLAS, camera at y = 1.5 with right = -x."""
    cam = CameraSpec(position=(0, 1.5, -5.5), forward=(0, 0, 1), right=(-1, 0, 0), up=(0, 1, 0), near=2.0)
    anim = GridAnimation()
    for _ in range(frame):
        anim.update()
    return SceneSpec("trippy_teapots", ["teapot.obj"], cam, anim.objects(), bench_size=(3840, 2160))


# ------------------------------------------------------------------------------------------ big_ben_clock (C5)
class BigBenAnimation:
    """examples/big_ben_clock.rs:67-103: twist every vertex about z, then ModelInstance::refit.
    `originals` are the BVH-REORDERED primitives (the example copies them after the model is built, :55-62)."""

    FRAC_2_PI = F(2.0 / np.pi)

    def __init__(self, originals):
        self.originals = np.ascontiguousarray(np.asarray(originals, F).reshape(-1, 3, 3)).copy()
        self.r = F(0)

    def animate(self):
        self.r = self.r + F(0.05)
        if self.r > self.FRAC_2_PI:
            self.r = self.r - self.FRAC_2_PI
        a = np.sin(self.r, dtype=F) * F(0.5)
        o = self.originals
        s = a * (o[:, :, 1] - F(0.2)) * F(0.2)
        c, sn = np.cos(s, dtype=F), np.sin(s, dtype=F)
        out = o.copy()
        out[:, :, 0] = o[:, :, 0] * c - o[:, :, 1] * sn
        out[:, :, 1] = o[:, :, 0] * sn + o[:, :, 1] * c
        return out.reshape(-1, 9)


def big_ben_clock():
    """examples/big_ben_clock.rs (C5): one instance, no with_transform (empty world bounds, identity)."""
    cam = CameraSpec(position=(0, 2.75, -2.5), forward=(0, 0, 1), right=(1, 0, 0), up=(0, 1, 0), near=2.0)
    return SceneSpec("big_ben_clock", ["bigben.tri"], cam, [ObjectSpec(0, with_transform=False)], bench_size=(7680, 4320))


def quad():
    """bvhtracer/tests/test_scene_quad.rs:52-131 (also examples/quad.rs): two triangles, BoxSpec camera."""
    cam = CameraSpec(position=(0, 0, 2), forward=(0, 0, -1), right=(1, 0, 0), up=(0, 1, 0), near=1.0,
                     fovy_deg=None, box=(-1, 1, -1, 1))
    return SceneSpec("quad", ["<quad>"], cam, [ObjectSpec(0)])


QUAD_TRIS = np.array([[-1, -1, 0, 1, 1, 0, -1, 1, 0], [-1, -1, 0, 1, -1, 0, 1, 1, 0]], F)
# examples/quad.rs:45-77: TextureCoordinates / Normals given to MeshBuilder::with_primitive, per triangle
QUAD_TEX_COORDS = np.array([[0, 0, 1, 1, 0, 1], [0, 0, 1, 0, 1, 1]], F)
QUAD_NORMALS = np.tile(np.array([0, 0, 1], F), (2, 3))


def quad_example(frame=0, elapsed=1.0 / 60.0):
    """examples/quad.rs:28-113: the textured quad, SymmetricFov camera at (0, 0, 2), spinning about z.

    The spin comes from the reference's rigid-body integrator (angular velocity (0, 0, 2), out of scope here); the
    stand-in is angle_z = 2 * elapsed * frame.  The texture (bricks_rgb.png, PNG decoding is host-side and out of
    scope) is replaced by `brick_texture` below."""
    cam = CameraSpec(position=(0, 0, 2), forward=(0, 0, -1), right=(1, 0, 0), up=(0, 1, 0), near=1.0)
    angle = float(F(2.0) * F(elapsed) * F(frame))
    return SceneSpec("quad_example", ["<quad>"], cam, [ObjectSpec(0, angle_z=angle)])


def brick_texture(width=256, height=256, seed=7):
    """Procedural Rgb<u8> stand-in for examples/assets/bricks_rgb.png: texels[height, width, 3]."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:height, 0:width]
    bh, bw = max(height // 8, 1), max(width // 4, 1)
    row = y // bh
    xs = x + (row % 2) * (bw // 2)
    mortar = ((y % bh) < max(bh // 8, 1)) | ((xs % bw) < max(bw // 16, 1))
    tone = rng.integers(-20, 20, (height // bh + 2, width // bw + 3))[row, xs // bw]
    tex = np.zeros((height, width, 3), np.int32)
    tex[..., 0] = 170 + tone; tex[..., 1] = 70 + tone // 2; tex[..., 2] = 50 + tone // 3
    tex[mortar] = (200, 200, 190)
    tex += rng.integers(-6, 7, tex.shape)
    return np.clip(tex, 0, 255).astype(np.uint8)

CONFIGS = {
    "cube": cube,
    "two_armadillos": two_armadillos,
    "sixteen_armadillos": sixteen_armadillos,
    "trippy_teapots": trippy_teapots,
    "big_ben_clock": big_ben_clock,
}
