"""Stub of osgeo (GDAL) for running the upstream LSsurf python in the build container.

TEST INFRASTRUCTURE ONLY (see oracle/README.md). The reference only touches gdal constants as
default-argument values at import time (fd_grid.py:147); mask/geotiff I/O is out of scope.
"""


class _Const:
    GDT_Float32 = 6
    GRA_NearestNeighbour = 0
    GRA_Average = 5

    def __getattr__(self, name):
        raise RuntimeError(f"osgeo stub: gdal.{name} is not available (mask/geotiff I/O is out of scope)")


gdal = _Const()
osr = _Const()
ogr = _Const()
