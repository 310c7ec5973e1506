"""sensor_msgs/PointCloud2 as plain data + the field mapping pcl::fromROSMsg applies for pcl::PointXYZI
(reference src/dlo/odom.cc:636-637).

ROS is not a dependency: `PointCloud2` mirrors the message's members one to one, so a rospy / rclpy message (or a
rosbag record) can be passed wherever one of these is expected — only the attributes below are read.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

# sensor_msgs/PointField datatypes
INT8, UINT8, INT16, UINT16, INT32, UINT32, FLOAT32, FLOAT64 = 1, 2, 3, 4, 5, 6, 7, 8
_SIZES = {INT8: 1, UINT8: 1, INT16: 2, UINT16: 2, INT32: 4, UINT32: 4, FLOAT32: 4, FLOAT64: 8}
_NP = {INT8: "i1", UINT8: "u1", INT16: "<i2", UINT16: "<u2", INT32: "<i4", UINT32: "<u4", FLOAT32: "<f4", FLOAT64: "<f8"}


@dataclass
class PointField:
    name: str
    offset: int
    datatype: int
    count: int = 1


@dataclass
class PointCloud2:
    height: int
    width: int
    fields: list = field(default_factory=list)
    is_bigendian: bool = False
    point_step: int = 0
    row_step: int = 0
    data: bytes = b""
    is_dense: bool = True


class Pc2Layout(C.Structure):
    """ngicp_pc2_layout (include/nanogicp_c.h)"""
    _fields_ = [("width", C.c_uint), ("height", C.c_uint), ("point_step", C.c_uint), ("row_step", C.c_uint),
                ("offset_x", C.c_int), ("offset_y", C.c_int), ("offset_z", C.c_int), ("offset_intensity", C.c_int),
                ("is_bigendian", C.c_int)]


def xyzi_layout(msg) -> Pc2Layout:
    """pcl::fromROSMsg's field mapping for pcl::PointXYZI (pcl::createMapping / FieldMatches): a point member is
    filled from the message field with the SAME NAME, datatype FLOAT32 and count 1 (count 0 is accepted as 1); members
    without such a field keep their default (intensity 0) and PCL prints "Failed to find match for field"."""
    off = {"x": -1, "y": -1, "z": -1, "intensity": -1}
    for f in msg.fields:
        if f.name in off and int(f.datatype) == FLOAT32 and int(f.count) in (0, 1) and off[f.name] < 0:
            off[f.name] = int(f.offset)
    return Pc2Layout(int(msg.width), int(msg.height), int(msg.point_step), int(msg.row_step),
                     off["x"], off["y"], off["z"], off["intensity"], 1 if msg.is_bigendian else 0)


def make_pointcloud2(xyzi: np.ndarray, kind: str = "ouster", height: int = 1, row_pad: int = 0, seed: int = 0) -> PointCloud2:
    """Synthetic messages in the layouts of the two common drivers (test/bench input; the other fields get noise).
    ouster: point_step 48 {x,y,z f32 @0,4,8; intensity f32 @16; t u32 @20; reflectivity u16 @24; ring u8 @26;
            ambient u16 @28; range u32 @32};  velodyne: point_step 22 {x,y,z,intensity f32 @0..12; ring u16 @16;
            time f32 @18} (unaligned records);  xyz: point_step 12, no intensity."""
    pts = np.asarray(xyzi, dtype=np.float32)
    n = pts.shape[0]
    assert n % height == 0
    width = n // height
    rng = np.random.default_rng(seed)
    if kind == "ouster":
        step = 48
        fields = [PointField("x", 0, FLOAT32), PointField("y", 4, FLOAT32), PointField("z", 8, FLOAT32), PointField("intensity", 16, FLOAT32),
                  PointField("t", 20, UINT32), PointField("reflectivity", 24, UINT16), PointField("ring", 26, UINT8),
                  PointField("ambient", 28, UINT16), PointField("range", 32, UINT32)]
    elif kind == "velodyne":
        step = 22
        fields = [PointField("x", 0, FLOAT32), PointField("y", 4, FLOAT32), PointField("z", 8, FLOAT32), PointField("intensity", 12, FLOAT32),
                  PointField("ring", 16, UINT16), PointField("time", 18, FLOAT32)]
    elif kind == "xyz":
        step = 12
        fields = [PointField("x", 0, FLOAT32), PointField("y", 4, FLOAT32), PointField("z", 8, FLOAT32)]
    else:
        raise ValueError(kind)
    row_step = width * step + row_pad
    buf = rng.integers(0, 256, size=height * row_step, dtype=np.uint8)
    rec = np.zeros((n, step), dtype=np.uint8)
    rec[:] = rng.integers(0, 256, size=(n, step), dtype=np.uint8)
    names = {"x": 0, "y": 1, "z": 2, "intensity": 4 if pts.shape[1] >= 5 else 3}
    for f in fields:
        if f.name in names and names[f.name] < pts.shape[1]:
            rec[:, f.offset:f.offset + 4] = np.ascontiguousarray(pts[:, names[f.name]]).view(np.uint8).reshape(n, 4)
    for r in range(height):
        buf[r * row_step: r * row_step + width * step] = rec[r * width:(r + 1) * width].reshape(-1)
    return PointCloud2(height=height, width=width, fields=fields, is_bigendian=False, point_step=step, row_step=row_step,
                       data=buf.tobytes(), is_dense=bool(np.isfinite(pts[:, :3]).all()))
