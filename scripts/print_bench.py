"""Prints the headline fields of a bench.py JSON line. usage: print_bench.py <file>"""
import json
import sys

d = json.load(open(sys.argv[1]))
print("value", round(d["value"], 2), d["unit"], "| e2e", round(d["e2e"]["value"], 2), "| roofline frac", round(d["roofline"]["frac"], 4),
      "| launches", d["gpu_launches"], "| clocks", d["clocks"])
print({k: round(d[k]["value"], 1) for k in ("png", "png_cfg4_shape", "bmp", "cfg5_shape") if k in d})
print("cpu_baseline", d["cpu_baseline"])
