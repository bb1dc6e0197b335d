"""DFMA rate of the probe for three operand patterns (irt_measure_fp64_rate): the register-file read limit of the
FP64 pipe that bounds K1's stage loop"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import irt_b200
ctx = irt_b200.Context(0)
names = {0: "a = fma(a, m, c)      m, c in the operand reuse cache     ",
         1: "a_k = fma(a_k, m, c_k)   two new register pairs per DFMA     ",
         2: "a_k = fma(a_k, b_k, c_k) three distinct register pairs / DFMA"}
base = None
for mode in (0, 1, 2):
    r = ctx.fp64_rate(mode) / 1e12
    base = base or r
    print("mode %d  %s  %6.2f TFLOP/s  (%.3f of mode 0)" % (mode, names[mode], r, r / base), flush=True)
