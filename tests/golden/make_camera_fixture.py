"""Extracts a handful of frames of the reference camera presets (cameras/rotate360.json, cameras/llff.json) into
tests/golden/cameras_fixture.json.  Run in the build container where /root/reference is mounted."""
import json
import os

REF = "/root/reference/cameras"
out = {}
r = json.load(open(os.path.join(REF, "rotate360.json")))
out["camera_angle_x"] = r["camera_angle_x"]
out["rotate360_total"] = len(r["frames"])
out["rotate360"] = {str(i): r["frames"][i]["transform_matrix"] for i in (0, 1, 45, 90, 180, 359, 540, 719)}
l = json.load(open(os.path.join(REF, "llff.json")))
out["llff_total"] = len(l["frames"])
out["llff"] = {str(i): l["frames"][i]["transform_matrix"] for i in range(0, len(l["frames"]), 50)}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cameras_fixture.json"), "w"), indent=1)
print("frames:", len(out["rotate360"]), len(out["llff"]))
