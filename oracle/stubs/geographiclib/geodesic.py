class _Unavailable:
    def Inverse(self, *args, **kwargs):
        raise RuntimeError("geographiclib is not installed in this image (stub from oracle/stubs)")


class Geodesic:
    WGS84 = _Unavailable()
