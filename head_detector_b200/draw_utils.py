"""Drawing helpers of the result object (reference: head_detector/draw_utils.py:15-90) - host-side cv2, same outputs:
`draw_bboxes` (blue-in-RGB 2 px rectangle), `draw_3d_landmarks` (landmark wireframe + head points), `draw_2d_landmarks`
(face points), `draw_pose` (three axis arrows from roll / pitch / yaw)."""
from math import cos, sin, sqrt

import cv2
import numpy as np

from .mesh import tables

POINT_COLOR = (255, 255, 255)


def draw_points(image: np.ndarray, points: np.ndarray, color=None) -> np.ndarray:
    radius = max(1, int(min(image.shape[:2]) * 0.001))
    for x, y in np.asarray(points)[:, :2]:
        cv2.circle(image, (int(x), int(y)), radius, POINT_COLOR if color is None else color, -1)
    return image


def draw_2d_landmarks(image: np.ndarray, head) -> np.ndarray:
    return draw_points(image, head.vertices_3d[tables()["face"], :2])


def draw_3d_landmarks(image: np.ndarray, head) -> np.ndarray:
    xy = head.vertices_3d[:, :2]
    for tri in xy[tables()["triangles"]].astype(np.int32):   # float -> int32 truncation as np.array(pts, np.int32) does
        cv2.polylines(image, [tri.reshape(-1, 1, 2)], isClosed=True, color=(0, 0, 255), thickness=1)
    return draw_points(image, xy[tables()["head_indices"]])


def draw_pose(image: np.ndarray, head) -> np.ndarray:
    rpy, box = head.head_pose, head.bbox
    area = box.w * box.h
    cx, cy = box.x + box.w // 2, box.y + box.h // 2
    size = sqrt(area) // 4
    pitch, yaw, roll = rpy.pitch * np.pi / 180, -(rpy.yaw * np.pi / 180), rpy.roll * np.pi / 180
    ends = ((size * (cos(yaw) * cos(roll)) + cx, size * (cos(pitch) * sin(roll) + cos(roll) * sin(pitch) * sin(yaw)) + cy, (0, 0, 255)),
            (size * (-cos(yaw) * sin(roll)) + cx, size * (cos(pitch) * cos(roll) - sin(pitch) * sin(yaw) * sin(roll)) + cy, (0, 255, 0)),
            (size * sin(yaw) + cx, size * (-cos(yaw) * sin(pitch)) + cy, (255, 0, 0)))
    thickness = max(1, int(sqrt(area) * 0.03))
    for x, y, color in ends:
        cv2.arrowedLine(image, (int(cx), int(cy)), (int(x), int(y)), color, thickness)
    return image


def draw_bboxes(image: np.ndarray, head) -> np.ndarray:
    x, y, w, h = (int(v) for v in head.bbox)
    cv2.rectangle(image, (x, y), (x + w, y + h), (255, 0, 0), 2)
    return image


DRAW_MAPPING = {"landmarks": [draw_3d_landmarks], "points": [draw_2d_landmarks], "pose": [draw_pose],
                "full": [draw_bboxes, draw_3d_landmarks], "bbox": [draw_bboxes]}
