"""B200-native message-passing engine behind the matdeeplearn.models operator
surface.  See DESIGN.md."""
__version__ = "0.1.0"
