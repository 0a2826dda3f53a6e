"""The uQ container: an uncompressed tar of NPY members (no .npy suffix) plus config.json.
Mirrors write_out / the archiving step (uq.py:273-274, 898-912) and the loader of the decoder
(uq.py:939-973).  Members are written in sorted name order and QNAME columns are read back in numeric
order (the reference's os.listdir order scrambles QNAME columns on decode).

The writer is a PLAN plus positioned writes: once the shape of every member is known, the byte
offset of every tar header, NPY header and payload follows (512-byte tar header, NPY v1.0 header padded to 64 bytes,
payload, padding to 512).  Payloads are then written straight from the buffers they were downloaded into - pinned
host memory filled by the device->host copies - with os.pwrite: no NPY blob, no BytesIO, no tarfile copy in between.
Because every slice has its own byte offset, the ranks of a multi-GPU encode write their slices of one container
side by side (multigpu.write_container_sharded) instead of gathering them on one rank.  The bytes of the file are
exactly those `tarfile` + `numpy.save` produce (tests/test_container_cli.py)."""
import io
import json
import os
import tarfile

import numpy
import numpy.lib.format as npf

BLOCK = tarfile.BLOCKSIZE          # 512
RECORD = tarfile.RECORDSIZE        # 10240: tarfile pads the archive to a multiple of this


def npy_bytes(arr):
    buf = io.BytesIO()
    numpy.save(buf, arr)                      # honours C / Fortran order exactly like the reference
    return buf.getvalue()


def npy_header(dtype, shape, fortran_order):
    """the header numpy.save writes in front of an array of this dtype / shape / memory order (NPY version 1.0)"""
    d = {'descr': npf.dtype_to_descr(numpy.dtype(dtype)), 'fortran_order': bool(fortran_order), 'shape': tuple(int(x) for x in shape)}
    buf = io.BytesIO()
    npf.write_array_header_1_0(buf, d)
    return buf.getvalue()


def npy_header_of(arr):
    """numpy.save(arr) header; like numpy, an array that is both C and Fortran contiguous is written as C"""
    fortran = arr.flags.f_contiguous and not arr.flags.c_contiguous
    return npy_header(arr.dtype, arr.shape, fortran)


def tar_header(name, size):
    info = tarfile.TarInfo(name)
    info.size = int(size)
    return info.tobuf(tarfile.DEFAULT_FORMAT, tarfile.ENCODING, "surrogateescape")


def payload_of(arr):
    """memory-order bytes of a C- or Fortran-contiguous array, without copying"""
    if arr.flags.c_contiguous:
        return arr.reshape(-1).view(numpy.uint8)
    if arr.flags.f_contiguous:
        return arr.T.reshape(-1).view(numpy.uint8)
    return numpy.ascontiguousarray(arr).reshape(-1).view(numpy.uint8)


class Plan:
    """Byte layout of a container.  entries: {name: (dtype, shape, fortran_order)}; config: the dict of config.json.
    offsets[name] = (offset of the tar header, offset of the payload, payload bytes); total = size of the file."""

    def __init__(self, entries, config):
        self.config_bytes = json.dumps(config, indent=4, sort_keys=True).encode()      # uq.py:903
        self.headers, self.offsets = {}, {}
        pos = 0
        names = sorted(list(entries) + ['config.json'])
        for name in names:
            if name == 'config.json':
                head, nbytes = b'', len(self.config_bytes)
            else:
                dtype, shape, fortran = entries[name]
                head = npy_header(dtype, shape, fortran)
                nbytes = int(numpy.prod(shape, dtype=numpy.int64)) * numpy.dtype(dtype).itemsize if len(shape) else numpy.dtype(dtype).itemsize
            size = len(head) + nbytes
            self.headers[name] = (pos, tar_header(name, size) + head)
            self.offsets[name] = (pos, pos + BLOCK + len(head), nbytes)
            pos += BLOCK + (size + BLOCK - 1) // BLOCK * BLOCK
        pos += 2 * BLOCK                                   # end-of-archive marker
        self.total = (pos + RECORD - 1) // RECORD * RECORD
        self.names = names

    def create(self, path):
        """write everything but the payloads: tar and NPY headers, config.json, zero padding, end blocks"""
        with open(path, 'wb') as f:
            f.truncate(self.total)                        # holes read as zeros: padding and end blocks are done
            for name in self.names:
                pos, head = self.headers[name]
                os.pwrite(f.fileno(), head, pos)
            os.pwrite(f.fileno(), self.config_bytes, self.offsets['config.json'][1])


def _pwrite_all(fd, view, offset):
    view = memoryview(view).cast('B')
    done = 0
    while done < len(view):
        done += os.pwrite(fd, view[done:done + (1 << 30)], offset + done)


def entries_of(members):
    return {name: (arr.dtype, arr.shape, arr.flags.f_contiguous and not arr.flags.c_contiguous) for name, arr in members.items()}


def write_container(path, members, config):
    """members: name -> ndarray exactly as the reference hands it to numpy.save (dtype, shape, memory order)"""
    plan = Plan(entries_of(members), config)
    plan.create(path)
    fd = os.open(path, os.O_WRONLY)
    try:
        for name, arr in members.items():
            _pwrite_all(fd, payload_of(arr), plan.offsets[name][1])
    finally:
        os.close(fd)
    return plan


def write_container_tarfile(path, members, config):
    """the same container through tarfile + numpy.save (three host copies per member); kept as the byte-for-byte
    reference of the planned writer in the tests"""
    blobs = {name: npy_bytes(arr) for name, arr in members.items()}
    blobs['config.json'] = json.dumps(config, indent=4, sort_keys=True).encode()
    with tarfile.open(path, mode='w') as tar:
        for name in sorted(blobs):
            info = tarfile.TarInfo(name)
            info.size = len(blobs[name])
            tar.addfile(info, io.BytesIO(blobs[name]))


def read_container(path_or_bytes, mmap=False):
    """-> (members: name -> ndarray as numpy.load returns it, config dict).  mmap=True (path only): members are
    numpy.memmap views of the file - nothing is read until the decoder uploads a member (or a rank its share of it)."""
    if isinstance(path_or_bytes, (bytes, bytearray)):
        tar = tarfile.open(fileobj=io.BytesIO(path_or_bytes))
        mmap = False
    else:
        if not tarfile.is_tarfile(path_or_bytes):                                      # uq.py:937
            raise ValueError('ERROR: Sorry, the path you have provided as input is a file, but not a tar file, and therefore cannot be a .uq file!')
        tar = tarfile.open(path_or_bytes)
    members, config = {}, None
    for info in tar.getmembers():
        name = info.name
        if name == 'config.json':
            config = json.loads(tar.extractfile(info).read().decode())
        elif mmap:
            f = tar.extractfile(info)
            version = npf.read_magic(f)
            shape, fortran, dtype = npf.read_array_header_1_0(f) if version == (1, 0) else npf.read_array_header_2_0(f)
            off = info.offset_data + f.tell()
            members[name] = numpy.memmap(path_or_bytes, dtype=dtype, mode='r', offset=off, shape=tuple(shape), order='F' if fortran else 'C')
        else:
            members[name] = numpy.load(io.BytesIO(tar.extractfile(info).read()))
    if config is None:                                                                 # uq.py:948-949
        raise ValueError('ERROR: No config.json file was found in your input path! I cannot decode data without it!')
    return members, config
