from .net5g import *
from .net5g_two_head import *
from .net6c import *
from .net6c_two_head import *
