// Pinned-host ring buffer feeding the GPU over a side stream (replaces the Petastorm reader
// pool + tf.data pipeline of the reference, P1/03:137-144,332-348).
//
//   producers (Python decode workers or the native synthetic/gather threads)
//        acquire_fill() -> write images/labels into the slot's pinned memory -> commit()
//   consumer (train loop)
//        next(): cudaMemcpyAsync(slot -> device buffer) on the copy stream, event-chained to the compute stream,
//                double-buffered on the device so the copy of batch k+1 overlaps the compute of batch k.
//
// Slots cycle FREE -> FILLING -> READY -> IN_FLIGHT(copy event) -> FREE.  All blocking waits release the GIL.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/extension.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace {

enum SlotState : int { FREE = 0, FILLING = 1, READY = 2, IN_FLIGHT = 3 };

struct Slot {
  uint8_t* img = nullptr;   // pinned
  int64_t* lab = nullptr;   // pinned
  SlotState state = FREE;
  cudaEvent_t copied = nullptr;
};

class RingLoader {
 public:
  RingLoader(int64_t num_slots, int64_t batch, int64_t image_bytes, int64_t device, int64_t device_bufs)
      : batch_(batch), image_bytes_(image_bytes), device_((int)device), use_cuda_(device >= 0) {
    TORCH_CHECK(num_slots >= 2 && batch >= 1 && image_bytes >= 1 && device_bufs >= 2);
    slots_.resize(num_slots);
    const size_t ibytes = (size_t)batch * image_bytes;
    for (auto& s : slots_) {
      if (use_cuda_) {
        C10_CUDA_CHECK(cudaHostAlloc((void**)&s.img, ibytes, cudaHostAllocPortable));
        C10_CUDA_CHECK(cudaHostAlloc((void**)&s.lab, (size_t)batch * sizeof(int64_t), cudaHostAllocPortable));
      } else {
        s.img = (uint8_t*)aligned_alloc(4096, (ibytes + 4095) / 4096 * 4096);
        s.lab = (int64_t*)aligned_alloc(4096, ((size_t)batch * 8 + 4095) / 4096 * 4096);
      }
    }
    if (use_cuda_) {
      c10::cuda::CUDAGuard guard(device_);
      for (auto& s : slots_) C10_CUDA_CHECK(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
      C10_CUDA_CHECK(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
      auto opt8 = at::TensorOptions().dtype(at::kByte).device(at::kCUDA, device_);
      auto opt64 = at::TensorOptions().dtype(at::kLong).device(at::kCUDA, device_);
      for (int i = 0; i < device_bufs; ++i) {
        dev_img_.push_back(at::empty({(int64_t)ibytes}, opt8));
        dev_lab_.push_back(at::empty({batch}, opt64));
        cudaEvent_t e;
        C10_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        consumed_.push_back(e);
        consumed_valid_.push_back(false);
      }
    }
  }
  ~RingLoader() {
    close();
    for (auto& s : slots_) {
      if (use_cuda_) {
        cudaFreeHost(s.img);
        cudaFreeHost(s.lab);
        if (s.copied) cudaEventDestroy(s.copied);
      } else {
        free(s.img);
        free(s.lab);
      }
    }
    for (auto e : consumed_) cudaEventDestroy(e);
    if (copy_stream_) cudaStreamDestroy(copy_stream_);
  }

  // ---------------------------------------------------------------- producer API
  int64_t acquire_fill() {
    pybind11::gil_scoped_release nogil;
    return acquire_fill_nogil();
  }
  std::pair<at::Tensor, at::Tensor> slot_tensors(int64_t slot) {
    TORCH_CHECK(slot >= 0 && slot < (int64_t)slots_.size());
    auto& s = slots_[slot];
    auto img = at::from_blob(s.img, {batch_ * image_bytes_}, at::TensorOptions().dtype(at::kByte));
    auto lab = at::from_blob(s.lab, {batch_}, at::TensorOptions().dtype(at::kLong));
    return {img, lab};
  }
  void commit(int64_t slot) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      TORCH_CHECK(slots_[slot].state == FILLING, "commit of a slot that is not being filled");
      slots_[slot].state = READY;
      ready_.push_back((int)slot);
    }
    cv_ready_.notify_one();
  }

  // ---------------------------------------------------------------- consumer API
  // returns (images uint8 [batch*image_bytes] on device, labels int64 [batch] on device)
  std::pair<at::Tensor, at::Tensor> next() {
    int slot = -1;
    {
      pybind11::gil_scoped_release nogil;
      std::unique_lock<std::mutex> lk(mu_);
      auto t0 = std::chrono::steady_clock::now();
      cv_ready_.wait(lk, [&] { return !ready_.empty() || closed_ || finished_; });
      wait_ns_ += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
      if (ready_.empty()) {
        // finite producers called finish(): everything committed has been handed out -> end of iteration
        if (finished_ && !closed_) throw pybind11::stop_iteration();
        throw std::runtime_error("RingLoader closed");
      }
      slot = ready_.front();
      ready_.pop_front();
      slots_[slot].state = IN_FLIGHT;
    }
    auto& s = slots_[slot];
    if (!use_cuda_) {
      auto img = at::empty({batch_ * image_bytes_}, at::kByte);
      auto lab = at::empty({batch_}, at::kLong);
      std::memcpy(img.data_ptr(), s.img, (size_t)batch_ * image_bytes_);
      std::memcpy(lab.data_ptr(), s.lab, (size_t)batch_ * 8);
      release_slot(slot);
      ++batches_;
      return {img, lab};
    }
    c10::cuda::CUDAGuard guard(device_);
    const int d = (int)(seq_ % dev_img_.size());
    cudaStream_t compute = at::cuda::getCurrentCUDAStream(device_);
    // the device buffer may still be read by the step that consumed it dev_bufs batches ago
    if (consumed_valid_[d]) C10_CUDA_CHECK(cudaStreamWaitEvent(copy_stream_, consumed_[d], 0));
    const size_t ibytes = (size_t)batch_ * image_bytes_;
    C10_CUDA_CHECK(cudaMemcpyAsync(dev_img_[d].data_ptr(), s.img, ibytes, cudaMemcpyHostToDevice, copy_stream_));
    C10_CUDA_CHECK(cudaMemcpyAsync(dev_lab_[d].data_ptr(), s.lab, (size_t)batch_ * 8, cudaMemcpyHostToDevice, copy_stream_));
    C10_CUDA_CHECK(cudaEventRecord(s.copied, copy_stream_));
    C10_CUDA_CHECK(cudaStreamWaitEvent(compute, s.copied, 0));
    // everything the compute stream has enqueued so far used the *previous* buffers: mark them consumed
    const int prev = (int)((seq_ + dev_img_.size() - 1) % dev_img_.size());
    C10_CUDA_CHECK(cudaEventRecord(consumed_[prev], compute));
    consumed_valid_[prev] = true;
    {
      std::lock_guard<std::mutex> lk(mu_);
      inflight_.push_back(slot);
    }
    ++seq_;
    ++batches_;
    h2d_bytes_ += ibytes + (size_t)batch_ * 8;
    return {dev_img_[d], dev_lab_[d]};
  }

  // ---------------------------------------------------------------- native fillers
  // Gather-collate threads: every batch is assembled from a pool of pre-generated "decoded" images
  // (uint8, JPEG-shaped H*W*3) - the collate cost of a real cached-dataset reader, without a network.
  void start_synthetic(int64_t threads, int64_t pool_images, int64_t num_classes, int64_t seed, int64_t shard,
                       int64_t num_shards) {
    TORCH_CHECK(workers_.empty(), "fillers already running");
    pool_.resize((size_t)pool_images * image_bytes_);
    pool_labels_.resize(pool_images);
    uint64_t x = 0x9E3779B97F4A7C15ull ^ (uint64_t)seed;
    uint64_t* p64 = reinterpret_cast<uint64_t*>(pool_.data());
    for (size_t i = 0; i < pool_.size() / 8; ++i) {
      x ^= x << 13; x ^= x >> 7; x ^= x << 17;
      p64[i] = x;
    }
    for (int64_t i = 0; i < pool_images; ++i) pool_labels_[i] = (int64_t)((i * 2654435761ull) % (uint64_t)num_classes);
    for (int t = 0; t < threads; ++t) {
      workers_.emplace_back([this, t, pool_images, seed, shard, num_shards] {
        uint64_t r = 0xD1B54A32D192ED03ull * (uint64_t)(t + 1) ^ (uint64_t)seed ^ ((uint64_t)shard << 32);
        while (true) {
          const int64_t slot = acquire_fill_nogil();
          if (slot < 0) return;
          auto& s = slots_[slot];
          for (int64_t b = 0; b < batch_; ++b) {
            r ^= r << 13; r ^= r >> 7; r ^= r << 17;
            // rows of this shard only: index = shard (mod num_shards)
            int64_t idx = (int64_t)(r % (uint64_t)pool_images);
            idx = idx - (idx % num_shards) + shard;
            if (idx >= pool_images) idx = shard;
            std::memcpy(s.img + b * image_bytes_, pool_.data() + idx * image_bytes_, image_bytes_);
            s.lab[b] = pool_labels_[idx];
          }
          commit_nogil(slot);
        }
      });
    }
  }

  // Producers are done (finite number of epochs): next() drains the committed batches, then raises StopIteration.
  void finish() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      finished_ = true;
    }
    cv_ready_.notify_all();
  }

  void close() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (closed_) return;
      closed_ = true;
    }
    cv_free_.notify_all();
    cv_ready_.notify_all();
    for (auto& w : workers_) w.join();
    workers_.clear();
    if (use_cuda_ && copy_stream_) cudaStreamSynchronize(copy_stream_);
  }

  int64_t batches() const { return batches_; }
  int64_t h2d_bytes() const { return h2d_bytes_; }
  double consumer_wait_ms() const { return wait_ns_ * 1e-6; }
  int64_t num_slots() const { return (int64_t)slots_.size(); }

 private:
  void release_slot(int slot) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      slots_[slot].state = FREE;
    }
    cv_free_.notify_one();
  }
  // recycle slots whose H2D copy has completed; caller holds mu_
  void reap_locked() {
    for (auto it = inflight_.begin(); it != inflight_.end();) {
      if (cudaEventQuery(slots_[*it].copied) == cudaSuccess) {
        slots_[*it].state = FREE;
        it = inflight_.erase(it);
      } else {
        ++it;
      }
    }
  }
  int64_t acquire_fill_nogil() {
    std::unique_lock<std::mutex> lk(mu_);
    while (true) {
      if (closed_) return -1;
      if (use_cuda_) reap_locked();
      for (size_t i = 0; i < slots_.size(); ++i) {
        if (slots_[i].state == FREE) {
          slots_[i].state = FILLING;
          return (int64_t)i;
        }
      }
      cv_free_.wait_for(lk, std::chrono::microseconds(200));
    }
  }
  void commit_nogil(int64_t slot) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      slots_[slot].state = READY;
      ready_.push_back((int)slot);
    }
    cv_ready_.notify_one();
  }

  int64_t batch_, image_bytes_;
  int device_;
  bool use_cuda_;
  std::vector<Slot> slots_;
  std::deque<int> ready_;
  std::deque<int> inflight_;
  std::mutex mu_;
  std::condition_variable cv_free_, cv_ready_;
  bool closed_ = false;
  bool finished_ = false;
  cudaStream_t copy_stream_ = nullptr;
  std::vector<at::Tensor> dev_img_, dev_lab_;
  std::vector<cudaEvent_t> consumed_;
  std::vector<bool> consumed_valid_;
  uint64_t seq_ = 0;
  std::atomic<int64_t> batches_{0};
  std::atomic<int64_t> h2d_bytes_{0};
  std::atomic<int64_t> wait_ns_{0};
  std::vector<std::thread> workers_;
  std::vector<uint8_t> pool_;
  std::vector<int64_t> pool_labels_;
};

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  namespace py = pybind11;
  m.doc() = "b200ddl pinned-host ring-buffer loader (side-stream H2D)";
  py::class_<RingLoader>(m, "RingLoader")
      .def(py::init<int64_t, int64_t, int64_t, int64_t, int64_t>(), py::arg("num_slots"), py::arg("batch"),
           py::arg("image_bytes"), py::arg("device"), py::arg("device_bufs") = 2)
      .def("acquire_fill", &RingLoader::acquire_fill)
      .def("slot_tensors", &RingLoader::slot_tensors)
      .def("commit", &RingLoader::commit)
      .def("next", &RingLoader::next)
      .def("start_synthetic", &RingLoader::start_synthetic, py::arg("threads"), py::arg("pool_images"),
           py::arg("num_classes"), py::arg("seed") = 0, py::arg("shard") = 0, py::arg("num_shards") = 1)
      .def("finish", &RingLoader::finish)
      .def("close", &RingLoader::close)
      .def_property_readonly("batches", &RingLoader::batches)
      .def_property_readonly("h2d_bytes", &RingLoader::h2d_bytes)
      .def_property_readonly("consumer_wait_ms", &RingLoader::consumer_wait_ms)
      .def_property_readonly("num_slots", &RingLoader::num_slots);
}
