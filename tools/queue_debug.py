import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sdrangel_b200 import DownChannelizerBank, capi
capi.init(0)
order = np.zeros(32, dtype=np.int32)
capi.check(capi.lib().b200dsp_probe_sm_order(32, 640, order.ctypes.data))
print("order", order.tolist())
rs = np.random.RandomState(1)
x = rs.randint(-32768, 32768, size=(200_000, 2)).astype(np.int16)
fs = 122880000
def run(rsv, n=200_000, chans=(60000, 180000)):
    b = DownChannelizerBank(fs)
    ids = [b.add_channel(48000, fc)[0] for fc in chans]
    if rsv is not None:
        b.set_reserved_sms(rsv)
    b.feed(x[:n])
    out = [b.fetch(i) for i in ids]
    b.close()
    return out
ref = run(None)
for rsv in ([], [147], [0], order[:8].tolist(), order[:24].tolist()):
    got = run(rsv)
    print(len(rsv), [bool(np.array_equal(a, g)) for a, g in zip(ref, got)], [int(np.abs(g).sum()) for g in got], [int(np.abs(a).sum()) for a in ref])
