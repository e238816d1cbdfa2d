"""Stand-in for diffusers 0.27.0 (oracle infrastructure, see ../README.md)."""
__version__ = "0.27.0-oracle-stub"
