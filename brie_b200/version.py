__version__ = "2.3.0+b200.r1"
