"""Value types shared by the host mirror: points, directions, point keys, constants."""
