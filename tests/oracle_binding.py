"""ctypes binding of the CPU oracle (oracle/liblz4ada_oracle.so).

Test infrastructure only.  Builds the oracle with its own Makefile when the
shared object is missing (gcc only, a second).
"""
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "liblz4ada_oracle.so")

EXC_NAMES = ["OK", "CHECKSUM_ERROR", "DATA_CORRUPTION", "NOT_SUPPORTED", "TOO_FEW_HEADER_BYTES",
             "TOO_LITTLE_MEMORY", "CONSTRAINT_ERROR", "ASSERTION_ERROR"]
RESERVATIONS = {"SZ_64_KiB": 0, "SZ_256_KiB": 1, "SZ_1_MiB": 2, "SZ_4_MiB": 3, "SZ_8_MiB": 4,
                "For_Modern": 3, "For_Legacy": 4, "For_All": 4, "Use_First": 5, "Single_Frame": 6}
EOF_NAMES = ["Yes", "No", "Maybe"]

c_u8p = ctypes.POINTER(ctypes.c_uint8)


class Xxh32State(ctypes.Structure):
    _fields_ = [("state", ctypes.c_uint32 * 4), ("buffer", ctypes.c_uint8 * 16),
                ("buffer_size", ctypes.c_int), ("total_length", ctypes.c_uint64)]


def _buf(b):
    return (ctypes.c_uint8 * max(1, len(b))).from_buffer_copy(bytes(b) if len(b) else b"\0")


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        L = lib
        L.lzo_init.restype = ctypes.c_void_p
        L.lzo_init.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        L.lzo_init_with_header.restype = ctypes.c_void_p
        L.lzo_init_with_header.argtypes = [c_u8p, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                           ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                                           ctypes.POINTER(ctypes.c_int), ctypes.c_char_p, ctypes.c_size_t]
        L.lzo_init_for_block.restype = ctypes.c_void_p
        L.lzo_init_for_block.argtypes = [ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_int]
        L.lzo_update.restype = ctypes.c_int
        L.lzo_update.argtypes = [ctypes.c_void_p, c_u8p, ctypes.c_int, ctypes.POINTER(ctypes.c_int), c_u8p,
                                 ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        L.lzo_is_end_of_frame.restype = ctypes.c_int
        L.lzo_is_end_of_frame.argtypes = [ctypes.c_void_p]
        L.lzo_message.restype = ctypes.c_char_p
        L.lzo_message.argtypes = [ctypes.c_void_p]
        L.lzo_free.argtypes = [ctypes.c_void_p]
        L.lzo_xxh32_hash.restype = ctypes.c_uint32
        L.lzo_xxh32_hash.argtypes = [c_u8p, ctypes.c_size_t]
        L.lzo_xxh32_reset.argtypes = [ctypes.POINTER(Xxh32State), ctypes.c_uint32]
        L.lzo_xxh32_update.argtypes = [ctypes.POINTER(Xxh32State), c_u8p, ctypes.c_size_t]
        L.lzo_xxh32_final.restype = ctypes.c_uint32
        L.lzo_xxh32_final.argtypes = [ctypes.POINTER(Xxh32State)]
        L.lzo_decode_stream.restype = ctypes.c_int
        L.lzo_decode_stream.argtypes = [c_u8p, ctypes.c_size_t, ctypes.c_size_t, c_u8p, ctypes.c_size_t,
                                        ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_int),
                                        ctypes.c_char_p, ctypes.c_size_t]
        L.lzo_decode_error_case.restype = ctypes.c_int
        L.lzo_decode_error_case.argtypes = [c_u8p, ctypes.c_size_t, c_u8p, ctypes.c_size_t,
                                            ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p, ctypes.c_size_t]

    # -- one-shot helpers -------------------------------------------------
    def xxh32(self, data):
        return self.lib.lzo_xxh32_hash(_buf(data), len(data))

    def decode_stream(self, data, chunk=4096, out_cap=None):
        """Init(For_All) + Update loop (lz4test.adb:32-83).  -> (exc_name, out, eof_name, msg)"""
        if out_cap is None:
            out_cap = max(1 << 16, 300 * len(data) + (1 << 16))
        out = (ctypes.c_uint8 * out_cap)()
        n = ctypes.c_size_t(0)
        eof = ctypes.c_int(0)
        msg = ctypes.create_string_buffer(700)
        rc = self.lib.lzo_decode_stream(_buf(data), len(data), chunk, out, out_cap, ctypes.byref(n),
                                        ctypes.byref(eof), msg, 700)
        return EXC_NAMES[rc], bytes(memoryview(out)[:n.value]), EOF_NAMES[eof.value], msg.value.decode()

    def decode_error_case(self, data, out_cap=1 << 24):
        """Init_With_Header(all, Single_Frame) + Update (lz4test.adb:280-308). -> (exc_name, out, msg)"""
        out = (ctypes.c_uint8 * out_cap)()
        n = ctypes.c_size_t(0)
        msg = ctypes.create_string_buffer(700)
        rc = self.lib.lzo_decode_error_case(_buf(data), len(data), out, out_cap, ctypes.byref(n), msg, 700)
        return EXC_NAMES[rc], bytes(memoryview(out)[:n.value]), msg.value.decode()

    # -- streaming API ----------------------------------------------------
    def init(self, reservation="For_All"):
        mb = ctypes.c_int(0)
        h = self.lib.lzo_init(RESERVATIONS[reservation], ctypes.byref(mb))
        return OracleCtx(self, h, mb.value)

    def init_with_header(self, data, reservation="Single_Frame"):
        nc, mb, exc = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
        msg = ctypes.create_string_buffer(700)
        h = self.lib.lzo_init_with_header(_buf(data), len(data), ctypes.byref(nc), ctypes.byref(mb),
                                          RESERVATIONS[reservation], ctypes.byref(exc), msg, 700)
        if not h:
            raise OracleError(EXC_NAMES[exc.value], msg.value.decode())
        return OracleCtx(self, h, mb.value), nc.value

    def init_for_block(self, compressed_length, reservation="For_All"):
        mb = ctypes.c_int(0)
        h = self.lib.lzo_init_for_block(ctypes.byref(mb), compressed_length, RESERVATIONS[reservation])
        return OracleCtx(self, h, mb.value)

    def hasher(self, seed=0):
        return OracleHasher(self, seed)


class OracleError(Exception):
    def __init__(self, name, message):
        super().__init__(message)
        self.name = name
        self.message = message


class OracleCtx:
    def __init__(self, oracle, handle, min_buffer_size):
        self.o, self.h, self.min_buffer_size = oracle, handle, min_buffer_size
        self.buffer = (ctypes.c_uint8 * min_buffer_size)()

    def update(self, data):
        """-> (num_consumed, output bytes, output_first, output_last)"""
        nc, of, ol = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
        rc = self.o.lib.lzo_update(self.h, _buf(data), len(data), ctypes.byref(nc), self.buffer,
                                   self.min_buffer_size, ctypes.byref(of), ctypes.byref(ol))
        if rc != 0:
            raise OracleError(EXC_NAMES[rc], self.o.lib.lzo_message(self.h).decode())
        out = bytes(memoryview(self.buffer)[of.value:ol.value + 1]) if ol.value >= of.value else b""
        return nc.value, out, of.value, ol.value

    def is_end_of_frame(self):
        return EOF_NAMES[self.o.lib.lzo_is_end_of_frame(self.h)]

    def __del__(self):
        if self.h:
            self.o.lib.lzo_free(self.h)
            self.h = None


class OracleHasher:
    def __init__(self, oracle, seed=0):
        self.o = oracle
        self.s = Xxh32State()
        oracle.lib.lzo_xxh32_reset(ctypes.byref(self.s), seed)

    def update(self, data):
        self.o.lib.lzo_xxh32_update(ctypes.byref(self.s), _buf(data), len(data))

    def final(self):
        return self.o.lib.lzo_xxh32_final(ctypes.byref(self.s))


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def load():
    src = os.path.join(ROOT, "oracle", "lz4ada_oracle.c")
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(src):
        build()
    return Oracle(ctypes.CDLL(SO))
