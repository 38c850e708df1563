from .boundary import *
from .collision import *
from .flows import *
from .reporter import *
from .vtk import *
from .bounce_back import *
from .hdf5 import *
