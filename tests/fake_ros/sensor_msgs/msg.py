class Image:
    def __init__(self, height=0, width=0, data=b""):
        self.height, self.width, self.data = height, width, data


class CompressedImage:
    def __init__(self, data=b""):
        self.data = data


class PointField:
    FLOAT32 = 7

    def __init__(self, name="", offset=0, datatype=7, count=1):
        self.name, self.offset, self.datatype, self.count = name, offset, datatype, count


class PointCloud2:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class PointCloud:
    pass
