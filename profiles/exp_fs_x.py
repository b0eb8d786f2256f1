import sys, os; sys.path.insert(0, ".")
from peakondrake_b200 import _lib
_lib.LIB_PATH = os.path.join(os.getcwd(), 'peakondrake_b200', 'libpeakon_b200_x.so')
exec(open("profiles/exp_fs.py").read())
