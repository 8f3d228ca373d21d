/* TEST INFRASTRUCTURE — shim `libaio.h` (the image has no libaio) for building the unmodified
 * reference (include/file_handles/flash_file_handle.h:6, CMakeLists.txt:112).  The five libaio
 * entry points the reference uses (src/file_handles/flash_file_handle.cpp:43,54,83,160,182) are thin
 * wrappers over the Linux native-AIO system calls, which is all libaio itself is; `struct iocb`
 * and `struct io_event` are the kernel ABI structures of <linux/aio_abi.h>. */
#pragma once
#include <errno.h>
#include <linux/aio_abi.h>
#include <string.h>
#include <sys/syscall.h>
#include <time.h>
#include <unistd.h>

typedef aio_context_t io_context_t;

/* libaio returns -errno instead of setting errno; the reference only compares against the request count */
static inline int io_setup(int maxevents, io_context_t* ctxp) {
  long r = syscall(SYS_io_setup, maxevents, ctxp);
  return r < 0 ? -errno : (int) r;
}
static inline int io_destroy(io_context_t ctx) {
  long r = syscall(SYS_io_destroy, ctx);
  return r < 0 ? -errno : (int) r;
}
static inline int io_submit(io_context_t ctx, long nr, struct iocb** ios) {
  long r = syscall(SYS_io_submit, ctx, nr, ios);
  return r < 0 ? -errno : (int) r;
}
static inline int io_getevents(io_context_t ctx, long min_nr, long nr, struct io_event* events,
                               struct timespec* timeout) {
  long r = syscall(SYS_io_getevents, ctx, min_nr, nr, events, timeout);
  return r < 0 ? -errno : (int) r;
}
static inline void io_prep_pread(struct iocb* cb, int fd, void* buf, size_t count, long long offset) {
  memset(cb, 0, sizeof(*cb));
  cb->aio_fildes = fd;
  cb->aio_lio_opcode = IOCB_CMD_PREAD;
  cb->aio_buf = (__u64) (unsigned long) buf;
  cb->aio_nbytes = count;
  cb->aio_offset = offset;
}
static inline void io_prep_pwrite(struct iocb* cb, int fd, void* buf, size_t count, long long offset) {
  memset(cb, 0, sizeof(*cb));
  cb->aio_fildes = fd;
  cb->aio_lio_opcode = IOCB_CMD_PWRITE;
  cb->aio_buf = (__u64) (unsigned long) buf;
  cb->aio_nbytes = count;
  cb->aio_offset = offset;
}
