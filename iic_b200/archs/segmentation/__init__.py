from .net10a import *
from .net10a_twohead import *
