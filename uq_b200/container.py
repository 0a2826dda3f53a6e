"""The uQ container: an uncompressed tar of NPY members (no .npy suffix) plus config.json.
Mirrors write_out / the archiving step (uq.py:273-274, 898-912) and the loader of the decoder
(uq.py:939-973).  Members are written in sorted name order and QNAME columns are read back in numeric
order (the reference's os.listdir order scrambles QNAME columns on decode)."""
import io
import json
import tarfile

import numpy


def npy_bytes(arr):
    buf = io.BytesIO()
    numpy.save(buf, arr)                      # honours C / Fortran order exactly like the reference
    return buf.getvalue()


def write_container(path, members, config):
    blobs = {name: npy_bytes(arr) for name, arr in members.items()}
    blobs['config.json'] = json.dumps(config, indent=4, sort_keys=True).encode()      # uq.py:903
    with tarfile.open(path, mode='w') as tar:
        for name in sorted(blobs):
            info = tarfile.TarInfo(name)
            info.size = len(blobs[name])
            tar.addfile(info, io.BytesIO(blobs[name]))


def read_container(path_or_bytes):
    """-> (members: name -> ndarray as numpy.load returns it, config dict)"""
    if isinstance(path_or_bytes, (bytes, bytearray)):
        tar = tarfile.open(fileobj=io.BytesIO(path_or_bytes))
    else:
        if not tarfile.is_tarfile(path_or_bytes):                                      # uq.py:937
            raise ValueError('ERROR: Sorry, the path you have provided as input is a file, but not a tar file, and therefore cannot be a .uq file!')
        tar = tarfile.open(path_or_bytes)
    members, config = {}, None
    for name in tar.getnames():
        data = tar.extractfile(name).read()
        if name == 'config.json':
            config = json.loads(data.decode())
        else:
            members[name] = numpy.load(io.BytesIO(data))
    if config is None:                                                                 # uq.py:948-949
        raise ValueError('ERROR: No config.json file was found in your input path! I cannot decode data without it!')
    return members, config
