#include "arrow_utils.hpp"

#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>

namespace {
struct ResultOwner {
  void *data = nullptr;
  size_t bytes = 0;
  bool shm = false;
  int fd = -1;
  const void *buffers[2] = {nullptr, nullptr};
  static constexpr const char *kShmName = "/warpdb_result";   // src/arrow_utils.cpp:45
};
void release_array(ArrowArray *a) {
  if (!a || !a->private_data) return;
  auto *o = static_cast<ResultOwner *>(a->private_data);
  if (o->shm) {
    munmap(o->data, o->bytes);
    if (o->fd >= 0) { close(o->fd); shm_unlink(ResultOwner::kShmName); }
  } else {
    std::free(o->data);
  }
  delete o;
  a->private_data = nullptr;
  a->release = nullptr;
}
void release_schema(ArrowSchema *s) { s->release = nullptr; }
}  // namespace

void export_to_arrow(const float *data, int64_t length, bool use_shared_memory, ArrowArray *out_array, ArrowSchema *out_schema) {
  if (!out_array || !out_schema) throw std::invalid_argument("Null output");
  auto *o = new ResultOwner();
  o->bytes = sizeof(float) * static_cast<size_t>(length);
  o->shm = use_shared_memory;
  if (use_shared_memory) {
    o->fd = shm_open(ResultOwner::kShmName, O_CREAT | O_RDWR, 0600);
    if (o->fd < 0) { delete o; throw std::runtime_error("shm_open failed"); }
    if (ftruncate(o->fd, static_cast<off_t>(o->bytes)) != 0) { close(o->fd); delete o; throw std::runtime_error("ftruncate failed"); }
    o->data = mmap(nullptr, o->bytes ? o->bytes : 1, PROT_READ | PROT_WRITE, MAP_SHARED, o->fd, 0);
    if (o->data == MAP_FAILED) { close(o->fd); delete o; throw std::runtime_error("mmap failed"); }
  } else {
    o->data = std::malloc(o->bytes ? o->bytes : 1);
    if (!o->data) { delete o; throw std::bad_alloc(); }
  }
  if (o->bytes) std::memcpy(o->data, data, o->bytes);
  o->buffers[1] = o->data;

  *out_array = ArrowArray{};
  out_array->length = length;
  out_array->n_buffers = 2;
  out_array->buffers = o->buffers;
  out_array->release = release_array;
  out_array->private_data = o;

  *out_schema = ArrowSchema{};
  out_schema->format = "f";
  out_schema->name = "result";
  out_schema->flags = ARROW_FLAG_NULLABLE;
  out_schema->release = release_schema;
}
