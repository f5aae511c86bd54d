"""DCNv2 module layer (forward + backward through the drop-in `_ext`), see dcn_v2.py."""
