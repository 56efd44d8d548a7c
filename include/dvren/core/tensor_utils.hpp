// dvren/core/tensor_utils.hpp -- hp_tensor views over host buffers (contiguous, row-major).
// Public interface of the reference (include/dvren/core/tensor_utils.hpp:13-31).
#pragma once

#include <array>
#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <vector>

#include "hotpath/hp.h"

namespace dvren {

hp_tensor MakeHostTensor(void* data, hp_dtype dtype, const std::vector<int64_t>& shape);
hp_tensor MakeHostTensor(void* data, hp_dtype dtype, std::initializer_list<int64_t> shape);

template <typename T, size_t N>
hp_tensor MakeHostTensor(T* data, hp_dtype dtype, const std::array<int64_t, N>& shape) {
    return MakeHostTensor(static_cast<void*>(data), dtype, std::vector<int64_t>(shape.begin(), shape.end()));
}

template <typename T>
hp_tensor MakeHostTensor(std::vector<T>& buffer, hp_dtype dtype, const std::vector<int64_t>& shape) {
    return MakeHostTensor(static_cast<void*>(buffer.data()), dtype, shape);
}

template <typename T>
hp_tensor MakeHostTensor(std::vector<T>& buffer, hp_dtype dtype, std::initializer_list<int64_t> shape) {
    return MakeHostTensor(static_cast<void*>(buffer.data()), dtype, std::vector<int64_t>(shape));
}

}  // namespace dvren
