// imgui.h -- the one ImGui type the graph model uses (node position), so that graph.hpp keeps the
// reference's member layout without the GUI toolkit.
#pragma once
struct ImVec2 {
    float x = 0.0f, y = 0.0f;
    constexpr ImVec2() = default;
    constexpr ImVec2(float x_, float y_) : x(x_), y(y_) {}
};
