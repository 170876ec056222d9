"""MJCF-subset reader.

Reads the subset of MJCF that the reference's converter emits
(ref:scripts/make_mjcf_from_robot_description.py:600-657,726-755,1060-1155) and that its test scene
uses (ref:test/test_resources/test_robot.xml): nested bodies with hinge/slide/free-less joints,
primitive geoms, sites, ``<replicate>``, cameras, default classes and rangefinder sensors.  It
produces the flat ``Model`` the engine consumes plus a kinematic tree for forward kinematics, so
parity scenes can be authored as MJCF without libmujoco.

Ordering follows the MuJoCo compiler: bodies depth-first pre-order (world = 0), geoms / sites
grouped by body in body order.  ``<replicate>`` names are ``name + sep + zero-padded index``.
"""
from __future__ import annotations

import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

from .model import (GEOM_BOX, GEOM_CAPSULE, GEOM_CYLINDER, GEOM_ELLIPSOID, GEOM_MESH, GEOM_PLANE,
                    GEOM_SPHERE, Model, Poses, axis_angle_quat, quat2mat, quat_mul)

_GEOM_TYPES = {
    "plane": GEOM_PLANE,
    "sphere": GEOM_SPHERE,
    "capsule": GEOM_CAPSULE,
    "ellipsoid": GEOM_ELLIPSOID,
    "cylinder": GEOM_CYLINDER,
    "box": GEOM_BOX,
    "mesh": GEOM_MESH,
}


def _floats(s, n=None, default=None):
    if s is None:
        return None if default is None else np.array(default, dtype=np.float64)
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None and v.shape[0] < n:
        v = np.concatenate([v, np.zeros(n - v.shape[0])])
    return v


@dataclass
class Frame:
    pos: np.ndarray
    quat: np.ndarray


@dataclass
class Joint:
    name: str
    type: str
    pos: np.ndarray
    axis: np.ndarray
    qposadr: int


@dataclass
class Body:
    name: str
    parent: int
    frame: Frame
    joints: list = field(default_factory=list)
    weldid: int = 0


@dataclass
class Site:
    name: str
    body: int
    frame: Frame


@dataclass
class Camera:
    name: str
    body: int
    frame: Frame
    fovy: float
    resolution: tuple


@dataclass
class GeomLocal:
    body: int
    frame: Frame


@dataclass
class Rangefinder:
    name: str
    site: int
    cutoff: float


@dataclass
class MjcfScene:
    model: Model
    bodies: list
    geoms: list
    sites: list
    cameras: list
    rangefinders: list
    nq: int
    extent: float = 1.0
    znear: float = 0.01
    zfar: float = 50.0

    # -- forward kinematics --------------------------------------------------------------------
    def body_poses(self, qpos=None):
        qpos = np.zeros(self.nq) if qpos is None else np.asarray(qpos, dtype=np.float64)
        xpos = np.zeros((len(self.bodies), 3))
        xquat = np.zeros((len(self.bodies), 4))
        xquat[0] = (1, 0, 0, 0)
        for b, body in enumerate(self.bodies):
            if b == 0:
                continue
            p = body.parent
            pm = quat2mat(xquat[p])
            pos = xpos[p] + pm @ body.frame.pos
            quat = quat_mul(xquat[p], body.frame.quat)
            for j in body.joints:
                q = qpos[j.qposadr]
                if j.type == "hinge":
                    m = quat2mat(quat)
                    anchor = pos + m @ j.pos
                    quat = quat_mul(quat, axis_angle_quat(j.axis, q))
                    pos = anchor - quat2mat(quat) @ j.pos
                elif j.type == "slide":
                    pos = pos + quat2mat(quat) @ (j.axis / np.linalg.norm(j.axis)) * q
                else:
                    raise ValueError(f"unsupported joint type {j.type}")
            xpos[b] = pos
            xquat[b] = quat / np.linalg.norm(quat)
        xmat = np.stack([quat2mat(q) for q in xquat])
        return xpos, xmat, xquat

    def poses(self, qpos=None) -> Poses:
        bpos, bmat, _ = self.body_poses(qpos)
        g = len(self.geoms)
        gx = np.zeros((g, 3))
        gm = np.zeros((g, 9))
        for i, gl in enumerate(self.geoms):
            gx[i] = bpos[gl.body] + bmat[gl.body] @ gl.frame.pos
            gm[i] = (bmat[gl.body] @ quat2mat(gl.frame.quat)).reshape(9)
        return Poses(gx, gm)

    def site_poses(self, qpos=None):
        bpos, bmat, _ = self.body_poses(qpos)
        sx = np.zeros((len(self.sites), 3))
        sm = np.zeros((len(self.sites), 3, 3))
        for i, s in enumerate(self.sites):
            sx[i] = bpos[s.body] + bmat[s.body] @ s.frame.pos
            sm[i] = bmat[s.body] @ quat2mat(s.frame.quat)
        return sx, sm

    def site_id(self, name):
        for i, s in enumerate(self.sites):
            if s.name == name:
                return i
        raise KeyError(name)

    def body_id(self, name):
        for i, b in enumerate(self.bodies):
            if b.name == name:
                return i
        raise KeyError(name)


class _Parser:
    def __init__(self, root: ET.Element):
        self.root = root
        comp = root.find("compiler")
        self.degree = True
        if comp is not None and comp.get("angle", "degree") == "radian":
            self.degree = False
        self.defaults: dict[str, dict[str, dict[str, str]]] = {"main": {}}
        for d in root.findall("default"):
            self._parse_default(d, "main", top=True)
        self.bodies: list[Body] = [Body("world", -1, Frame(np.zeros(3), np.array([1.0, 0, 0, 0])))]
        self.body_geoms: dict[int, list] = {}
        self.body_sites: dict[int, list] = {}
        self.body_cams: dict[int, list] = {}
        self.nq = 0
        self.replicated_sites: dict[str, list[str]] = {}

    def _parse_default(self, el, parent_cls, top=False):
        cls = el.get("class", "main" if top else None)
        if cls is None:
            cls = "main"
        base = {k: dict(v) for k, v in self.defaults.get(parent_cls, {}).items()}
        for child in el:
            if child.tag == "default":
                continue
            base.setdefault(child.tag, {}).update(child.attrib)
        self.defaults[cls] = base
        for child in el.findall("default"):
            self._parse_default(child, cls)

    def _attrs(self, el, childclass=None):
        cls = el.get("class", childclass or "main")
        merged = dict(self.defaults.get(cls, self.defaults["main"]).get(el.tag, {}))
        merged.update(el.attrib)
        return merged

    def _frame(self, a) -> Frame:
        pos = _floats(a.get("pos"), 3, (0, 0, 0))
        if "quat" in a:
            quat = _floats(a["quat"])
            quat = quat / np.linalg.norm(quat)
        elif "euler" in a:
            quat = self._euler(_floats(a["euler"], 3))
        else:
            quat = np.array([1.0, 0, 0, 0])
        return Frame(pos, quat)

    def _euler(self, e) -> np.ndarray:
        if self.degree:
            e = np.deg2rad(e)
        q = np.array([1.0, 0, 0, 0])
        for ax, ang in zip(((1, 0, 0), (0, 1, 0), (0, 0, 1)), e):  # intrinsic xyz (MuJoCo default)
            q = quat_mul(q, axis_angle_quat(ax, ang))
        return q

    @staticmethod
    def _compose(outer: Frame, inner: Frame) -> Frame:
        return Frame(outer.pos + quat2mat(outer.quat) @ inner.pos, quat_mul(outer.quat, inner.quat))

    def _walk(self, el, body_id, pre: Frame, suffix: str, childclass=None):
        """Process children of ``el`` (a body / replicate / frame) attached to ``body_id``;
        ``pre`` is the accumulated replicate/frame transform inside that body."""
        for child in el:
            tag = child.tag
            if tag == "body":
                a = child.attrib
                fr = self._compose(pre, self._frame(a))
                bid = len(self.bodies)
                self.bodies.append(Body(a.get("name", f"body{bid}") + suffix, body_id, fr))
                self._walk(child, bid, Frame(np.zeros(3), np.array([1.0, 0, 0, 0])), suffix,
                           a.get("childclass", childclass))
            elif tag == "joint":
                a = self._attrs(child, childclass)
                jt = a.get("type", "hinge")
                j = Joint(a.get("name", "") + suffix, jt, _floats(a.get("pos"), 3, (0, 0, 0)),
                          _floats(a.get("axis"), 3, (0, 0, 1)), self.nq)
                self.nq += 1
                self.bodies[body_id].joints.append(j)
            elif tag == "geom":
                a = self._attrs(child, childclass)
                self.body_geoms.setdefault(body_id, []).append(
                    (a, self._compose(pre, self._frame(a)), suffix))
            elif tag == "site":
                a = self._attrs(child, childclass)
                self.body_sites.setdefault(body_id, []).append(
                    (a.get("name", "") + suffix, self._compose(pre, self._frame(a))))
                if suffix:
                    self.replicated_sites.setdefault(a.get("name", ""), []).append(
                        a.get("name", "") + suffix)
            elif tag == "camera":
                a = self._attrs(child, childclass)
                res = tuple(int(x) for x in a.get("resolution", "1 1").split())
                self.body_cams.setdefault(body_id, []).append(
                    (a.get("name", "") + suffix, self._compose(pre, self._frame(a)),
                     float(a.get("fovy", 45.0)), res))
            elif tag == "replicate":
                count = int(child.get("count"))
                sep = child.get("sep", "")
                offset = _floats(child.get("offset"), 3, (0, 0, 0))
                step = Frame(offset, self._euler(_floats(child.get("euler"), 3, (0, 0, 0))))
                width = len(str(count - 1))
                acc = Frame(np.zeros(3), np.array([1.0, 0, 0, 0]))
                for i in range(count):
                    self._walk(child, body_id, self._compose(pre, acc),
                               suffix + sep + str(i).zfill(width), childclass)
                    acc = self._compose(acc, step)
            elif tag == "frame":
                self._walk(child, body_id, self._compose(pre, self._frame(child.attrib)), suffix,
                           childclass)
            # inertial, light, ... : irrelevant to the sensor path

    def parse(self) -> MjcfScene:
        wb = self.root.find("worldbody")
        self._walk(wb, 0, Frame(np.zeros(3), np.array([1.0, 0, 0, 0])), "")
        # weld ids
        for b, body in enumerate(self.bodies):
            if b == 0:
                body.weldid = 0
            else:
                body.weldid = b if body.joints else self.bodies[body.parent].weldid
        gtype, gbody, ggroup, gsize, grgba, geoms = [], [], [], [], [], []
        for b in range(len(self.bodies)):
            for a, fr, _ in self.body_geoms.get(b, []):
                t = _GEOM_TYPES[a.get("type", "sphere")]
                if t == GEOM_MESH and "mesh" not in a:
                    raise ValueError("mesh geom without mesh asset")
                gtype.append(t)
                gbody.append(b)
                ggroup.append(int(a.get("group", 0)))
                gsize.append(_floats(a.get("size"), 3, (0, 0, 0)))
                grgba.append(_floats(a.get("rgba"), 4, (0.5, 0.5, 0.5, 1.0)))
                geoms.append(GeomLocal(b, fr))
        sites, cams = [], []
        for b in range(len(self.bodies)):
            for name, fr in self.body_sites.get(b, []):
                sites.append(Site(name, b, fr))
            for name, fr, fovy, res in self.body_cams.get(b, []):
                cams.append(Camera(name, b, fr, fovy, res))
        site_names = {s.name: i for i, s in enumerate(sites)}
        rfs = []
        sens = self.root.find("sensor")
        if sens is not None:
            for s in sens.findall("rangefinder"):
                base, sname = s.get("site"), s.get("name", "")
                cutoff = float(s.get("cutoff", 0.0))
                if base in site_names:
                    rfs.append(Rangefinder(sname, site_names[base], cutoff))
                else:  # replicated site: one sensor per replica, same suffix
                    for full in self.replicated_sites.get(base, []):
                        rfs.append(Rangefinder(sname + full[len(base):], site_names[full], cutoff))
        g = len(gtype)
        model = Model(
            geom_type=gtype, geom_bodyid=gbody, geom_group=ggroup, geom_matid=-np.ones(g, np.int32),
            geom_dataid=-np.ones(g, np.int32), geom_size=np.array(gsize).reshape(g, 3),
            geom_rgba=np.array(grgba, dtype=np.float32).reshape(g, 4),
            mat_rgba=np.zeros((0, 4), np.float32),
            body_weldid=[b.weldid for b in self.bodies])
        scene = MjcfScene(model, self.bodies, geoms, sites, cams, rfs, self.nq)
        stat = self.root.find("statistic")
        if stat is not None and stat.get("extent"):
            scene.extent = float(stat.get("extent"))
        vis = self.root.find("visual")
        if vis is not None and vis.find("map") is not None:
            mp = vis.find("map")
            scene.znear = float(mp.get("znear", scene.znear))
            scene.zfar = float(mp.get("zfar", scene.zfar))
        return scene


def load_mjcf_string(xml: str) -> MjcfScene:
    return _Parser(ET.fromstring(xml)).parse()


def load_mjcf(path: str) -> MjcfScene:
    return _Parser(ET.parse(path).getroot()).parse()
