// dvren/core/status.hpp -- error value of the dvren C++ surface.
// Same public interface as the reference (include/dvren/core/status.hpp:9-36,
// src/core/status.cpp:9-79): StatusCode mirrors hp_status one to one.
#pragma once

#include <string>

#include "hotpath/hp.h"

namespace dvren {

enum class StatusCode { kOk = 0, kInvalidArgument, kOutOfMemory, kNotImplemented, kUnsupported, kInternalError };

class Status {
public:
    Status() = default;
    Status(StatusCode code, std::string message) : code_(code), message_(std::move(message)) {}

    static Status Ok() { return {}; }
    static Status FromHotpath(hp_status code, std::string message = {});

    [[nodiscard]] bool ok() const { return code_ == StatusCode::kOk; }
    [[nodiscard]] explicit operator bool() const { return ok(); }
    [[nodiscard]] StatusCode code() const { return code_; }
    [[nodiscard]] const std::string& message() const { return message_; }
    // "ok", "<code name>" or "<code name>: <message>"
    [[nodiscard]] std::string ToString() const;

private:
    StatusCode code_{StatusCode::kOk};
    std::string message_{};
};

}  // namespace dvren
