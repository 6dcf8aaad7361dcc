"""Version of the reference interface this build mirrors (core/_version.py) with a local build tag."""
__version_info__ = (1, 22, 2)
__version__ = ".".join(str(v) for v in __version_info__) + "+b200"
