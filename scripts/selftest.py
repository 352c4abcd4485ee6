import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import probav_b200 as pb
from probav_b200 import _lib
import ctypes as C
lib=_lib.lib()
buf=C.create_string_buffer(1<<16)
try:
    rc=lib.pv_selftest(buf, len(buf))
    print("rc",rc); print(buf.value.decode())
except Exception as e:
    print("EXC", e)
