// oracle/ref_recipe/hostcl/CL/cl.hpp -- TEST INFRASTRUCTURE ONLY.
//
// A host-memory stand-in for the slice of Khronos' cl.hpp (OpenCL 1.2 C++ bindings) that the
// reference's waveguide::run template, its program classes and its stock pre/post-processors touch,
// so that those can be compiled UNMODIFIED and run here, where there is no OpenCL platform:
//   cl::Buffer        a reference-counted block of host memory (copies share it, as cl_mem handles do)
//   cl::CommandQueue  in order and synchronous: enqueueReadBuffer / enqueueWriteBuffer are memcpy
//   cl::copy          both directions
//   cl::make_kernel   looks the kernel up BY NAME in a registry of host functions; the functions
//                     registered are the reference's own OpenCL-C kernels as compiled for the host by
//                     the same recipe (ref_wg_f32.cpp), launched over the NDRange one work-item at a
//                     time. Buffers are passed as their pointers, everything else by value.
//   cl::Program / Context / Device   carry nothing: there is nothing to build or to choose.
// No arithmetic happens in this file.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iterator>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

typedef uint64_t cl_mem_flags;
typedef uint32_t cl_bool;
typedef uint32_t cl_program_info;
typedef intptr_t cl_context_properties;
#define CL_TRUE 1
#define CL_FALSE 0
#define CL_MEM_READ_WRITE (1 << 0)
#define CL_MEM_WRITE_ONLY (1 << 1)
#define CL_MEM_READ_ONLY (1 << 2)
#define CL_MEM_SIZE 0x1102
#define CL_SUCCESS 0

namespace cl {

class Context final {};
class Device final {};
class Program final {
public:
    template <cl_program_info Info>
    std::string getInfo() const { return {}; }   // nothing was built: no log, no binaries
};
class Event final {};

class Buffer final {
public:
    Buffer() = default;
    Buffer(const Context&, cl_mem_flags, size_t bytes)
            : mem_{std::make_shared<std::vector<unsigned char>>(bytes)} {}
    template <typename It>
    Buffer(const Context&, It begin, It end, bool /*read_only*/, bool /*use_host_ptr*/ = false) {
        using T = typename std::iterator_traits<It>::value_type;
        const size_t n = size_t(std::distance(begin, end));
        mem_ = std::make_shared<std::vector<unsigned char>>(n * sizeof(T));
        unsigned char* p = mem_->data();
        for (; begin != end; ++begin, p += sizeof(T)) {
            const T v = *begin;
            std::memcpy(p, &v, sizeof(T));
        }
    }
    template <int Info>
    size_t getInfo() const {
        static_assert(Info == CL_MEM_SIZE, "only CL_MEM_SIZE is answered");
        return bytes();
    }
    size_t bytes() const { return mem_ ? mem_->size() : 0; }
    unsigned char* data() const { return mem_ && !mem_->empty() ? mem_->data() : nullptr; }

private:
    std::shared_ptr<std::vector<unsigned char>> mem_;
};

class CommandQueue final {
public:
    CommandQueue() = default;
    CommandQueue(const Context&, const Device&) {}
    int enqueueReadBuffer(const Buffer& b, cl_bool, size_t offset, size_t size, void* ptr) const {
        if (offset + size > b.bytes()) throw std::runtime_error{"enqueueReadBuffer outside the buffer"};
        std::memcpy(ptr, b.data() + offset, size);
        return CL_SUCCESS;
    }
    int enqueueWriteBuffer(const Buffer& b, cl_bool, size_t offset, size_t size, const void* ptr) const {
        if (offset + size > b.bytes()) throw std::runtime_error{"enqueueWriteBuffer outside the buffer"};
        std::memcpy(b.data() + offset, ptr, size);
        return CL_SUCCESS;
    }
    int finish() const { return CL_SUCCESS; }
};

template <typename It>
int copy(const CommandQueue&, It begin, It end, Buffer& b) {
    using T = typename std::iterator_traits<It>::value_type;
    if (size_t(std::distance(begin, end)) * sizeof(T) > b.bytes()) throw std::runtime_error{"cl::copy outside the buffer"};
    unsigned char* p = b.data();
    for (; begin != end; ++begin, p += sizeof(T)) {
        const T v = *begin;
        std::memcpy(p, &v, sizeof(T));
    }
    return CL_SUCCESS;
}
template <typename It>
int copy(const CommandQueue&, const Buffer& b, It begin, It end) {
    using T = typename std::iterator_traits<It>::value_type;
    if (size_t(std::distance(begin, end)) * sizeof(T) > b.bytes()) throw std::runtime_error{"cl::copy outside the buffer"};
    const unsigned char* p = b.data();
    for (; begin != end; ++begin, p += sizeof(T)) {
        T v;
        std::memcpy(&v, p, sizeof(T));
        *begin = v;
    }
    return CL_SUCCESS;
}

class NDRange final {
public:
    NDRange(size_t n) : n{n} {}
    size_t n;
};

class EnqueueArgs final {
public:
    EnqueueArgs(CommandQueue queue, NDRange global) : queue{queue}, global{global} {}
    CommandQueue queue;
    NDRange global;
};

// name -> launcher(global size, argument pointers)
using hostcl_launcher = std::function<void(size_t, void**)>;
inline std::map<std::string, hostcl_launcher>& hostcl_registry() {
    static std::map<std::string, hostcl_launcher> registry;
    return registry;
}

namespace detail {
template <typename T>
void* hostcl_argument(T& by_value) { return &by_value; }
inline void* hostcl_argument(Buffer& b) { return b.data(); }
}  // namespace detail

template <typename... Ts>
class make_kernel final {
public:
    make_kernel(const Program&, const std::string& name, int* error = nullptr) : name_{name} {
        if (error) *error = CL_SUCCESS;
    }
    Event operator()(const EnqueueArgs& args, Ts... ts) const {
        const auto it = hostcl_registry().find(name_);
        if (it == hostcl_registry().end()) throw std::runtime_error{"no host kernel registered as " + name_};
        void* argv[] = {detail::hostcl_argument(ts)...};
        it->second(args.global.n, argv);
        return Event{};
    }

private:
    std::string name_;
};

}  // namespace cl
