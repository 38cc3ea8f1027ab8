"""Print the tcgen05 conv launch plan for a shape (host only, no GPU): python tools/tc_plan.py B Cin Cout X Y Z kx ky kz sx sy sz tr [nsplit]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vs_seg_b200 import lib as L


def describe(B, cin, cout, dims, k, s, tr, nsplit=0, sc=None):
    lib = L.load()
    X, Y, Z = dims
    if tr:
        od = (X * s[0], Y * s[1], Z * s[2])
    else:
        od = tuple((d + q - 1) // q for d, q in zip(dims, s))
    def act(Cc, d):
        n = d[0] * d[1] * d[2]
        return L.Act8(4096, B * Cc * n, Cc * n, B, Cc, *d)
    a, o = act(cin, dims), act(cout, od)
    g = L.ConvGeom(*k, *s, 1 if tr else 0)
    scv = act(sc, dims) if sc else None
    scp = C.byref(scv) if scv is not None else None
    ns = nsplit or lib.vsseg_conv3d_tc_suggest_split(C.byref(a), C.byref(o), C.byref(g), scp)
    buf = C.create_string_buffer(8192)
    if ns > 0:
        lib.vsseg_conv3d_tc_describe(C.byref(a), C.byref(o), C.byref(g), ns, scp, buf, 8192)
    return ns, buf.value.decode()


if __name__ == "__main__":
    v = [int(t) for t in sys.argv[1:]]
    print(describe(v[0], v[1], v[2], v[3:6], v[6:9], v[9:12], v[12], v[13] if len(v) > 13 else 0))
