"""Kernel list of a batch of one (for ncu): usage trace_one.py gz|png"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import debigulator_b200 as dbg
gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
ctx = dbg.Context(0)
if sys.argv[1] == "gz":
    gz = open(os.path.join(gold, "gzipsample.gz"), "rb").read()
    for _ in range(3): r = ctx.decode_gz_batch([gz], [600000])
else:
    png = open(os.path.join(gold, "gimp_test.png"), "rb").read()
    for _ in range(3): r = ctx.decode_png_batch([png])
print(r[0][0], ctx.lane_stats(), ctx.bsplit_stats())
