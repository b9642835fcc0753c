// THROW_LOGIC_ERROR(fmt, ...): internal-fault reporting with the source location prepended.
// Same call shape as the reference's include/utility/logic-error-utility.hpp:2-12.
#pragma once
#include <format>
#include <source_location>
#include <stdexcept>

#define THROW_LOGIC_ERROR(...)                                                                  \
    {                                                                                           \
        const auto nodey_loc_ = std::source_location::current();                                \
        throw std::logic_error(std::format("[{}:{}] {}", nodey_loc_.file_name(),                \
                                           nodey_loc_.line(), std::format(__VA_ARGS__)));       \
    }
