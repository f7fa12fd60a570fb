// chrono_b200.hpp -- header-only C++ host mirror of chrono-photo's processor / option interface on top of the C ABI
// (include/chrono_b200.h). The reference is a Rust crate; with no Rust toolchain in the build image this is the
// compiled-language host side: same type names, same string grammars, same error behaviour (exceptions where the
// reference returns Err / panics). Citations are into mlange-42/chrono-photo v0.6.5.
//
//   Threshold, Fade, FadeMode, BackgroundMode, OutlierSelectionMode, SelectionMode   src/options.rs
//   FrameRange                                                                        src/flist.rs:9-79
//   OutlierProcessor (new / process)                                                  src/chrono.rs:46-54, :73-81
//   SimpleProcessor (new / process)                                                   src/simple.rs:18, :26-32
//   GpuStack                                 replaces TimeSlicer::write_time_slices   src/slicer.rs:106
//   ShakeParams, ShakeAnchor, ShakeAnalyzer                                          src/shake.rs:44-119, :185-305
#pragma once
#include <cstdint>
#include <cstdlib>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "chrono_b200.h"

namespace chrono_b200 {

struct ParseEnumError : std::invalid_argument { using std::invalid_argument::invalid_argument; };    // src/lib.rs
struct ParseOptionError : std::invalid_argument { using std::invalid_argument::invalid_argument; };  // src/lib.rs
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) {
    if (rc != CHB_OK) throw Error(rc, chb_last_error());
}
inline std::vector<std::string> split(const std::string& s, char sep) {
    std::vector<std::string> out;
    std::string cur;
    for (char ch : s) {
        if (ch == sep) { out.push_back(cur); cur.clear(); } else cur.push_back(ch);
    }
    out.push_back(cur);
    return out;
}
inline float parse_f32(const std::string& s, const std::string& what) {
    char* end = nullptr;
    float v = std::strtof(s.c_str(), &end);
    if (s.empty() || end == nullptr || *end != '\0') throw ParseOptionError(what);
    return v;
}
inline int parse_i32(const std::string& s, const std::string& what) {
    char* end = nullptr;
    long v = std::strtol(s.c_str(), &end, 10);
    if (s.empty() || end == nullptr || *end != '\0') throw ParseOptionError(what);
    return (int)v;
}

enum class SelectionMode { Outlier, Lighter, Darker };  // src/options.rs:8-31
inline SelectionMode parse_selection_mode(const std::string& s) {
    if (s == "lighter") return SelectionMode::Lighter;
    if (s == "darker") return SelectionMode::Darker;
    if (s == "outlier") return SelectionMode::Outlier;
    throw ParseEnumError("Not a pixel selection mode: " + s + ". Must be one of (lighter|darker|outlier)");
}
enum class BackgroundMode { First = 0, Random = 1, Average = 2, Median = 3 };  // src/options.rs:319-344
inline BackgroundMode parse_background_mode(const std::string& s) {
    if (s == "first") return BackgroundMode::First;
    if (s == "random") return BackgroundMode::Random;
    if (s == "average") return BackgroundMode::Average;
    if (s == "median") return BackgroundMode::Median;
    throw ParseEnumError("Not a background pixel selection mode: " + s + ". Must be one of (first|random|average|median)");
}
enum class OutlierSelectionMode { First = 0, Last = 1, Extreme = 2, Average = 3, AllForward = 4, AllBackward = 5 };  // :284-315
inline OutlierSelectionMode parse_outlier_mode(const std::string& s) {
    if (s == "first") return OutlierSelectionMode::First;
    if (s == "last") return OutlierSelectionMode::Last;
    if (s == "extreme") return OutlierSelectionMode::Extreme;
    if (s == "average") return OutlierSelectionMode::Average;
    if (s == "forward") return OutlierSelectionMode::AllForward;
    if (s == "backward") return OutlierSelectionMode::AllBackward;
    throw ParseEnumError("Not an outlier selection mode: " + s + ". Must be one of (first|last|extreme|average|forward|backward)");
}
enum class FadeMode { Clamp = 0, Repeat = 1 };  // src/options.rs:35-56

// src/options.rs:188-280
class Threshold {
public:
    Threshold(bool absolute, float min, float max) : absolute_(absolute) { chb_threshold_new(absolute ? 1 : 0, min, max, &min_, &max_, &scale_); }
    static Threshold abs(float min, float max) { return Threshold(true, min, max); }
    static Threshold rel(float min, float max) { return Threshold(false, min, max); }
    static Threshold from_str(const std::string& s) {
        auto parts = split(s, '/');
        bool absolute;
        if (parts[0] == "absolute" || parts[0] == "abs") absolute = true;
        else if (parts[0] == "relative" || parts[0] == "rel") absolute = false;
        else throw ParseOptionError("Not a pixel outlier detection mode: " + s + ". Must be one of (abs[olute]|rel[ative])/<min>[/<max>]");
        if (parts.size() < 2) throw ParseOptionError("Unexpected format in " + s);
        float mn = parse_f32(parts[1], "Unable to parse lower threshold for outlier detection: " + s);
        float mx = parts.size() > 2 ? parse_f32(parts[2], "Unable to parse upper threshold for outlier detection: " + s) : mn;
        return Threshold(absolute, mn, mx);
    }
    bool absolute() const { return absolute_; }
    float min() const { return min_; }
    float max() const { return max_; }
    float scale() const { return scale_; }

private:
    bool absolute_;
    float min_, max_, scale_;
};

// src/options.rs:59-185
class Fade {
public:
    Fade(FadeMode mode, bool absolute, const std::vector<std::pair<int, float>>& frames) : is_none_(false), mode_(mode), absolute_(absolute) {
        if (frames.size() < 2) throw ParseOptionError("Fade requires at least two frames specified.");
        std::vector<int32_t> fr;
        std::vector<float> va;
        for (auto& p : frames) { fr.push_back(p.first); va.push_back(p.second); }
        const int cap = fr.back() - fr.front() + 1;
        if (cap < 1) throw ParseOptionError("Fade frames must be ordered by frame");
        values_.assign((size_t)cap, 1.0f);
        int n = chb_fade_build(fr.data(), va.data(), (int)fr.size(), values_.data(), cap, &offset_);
        if (n < 0) throw ParseOptionError("Invalid fade specification");
        values_.resize((size_t)n);
    }
    static Fade none() { return Fade(); }
    static Fade from_str(const std::string& s) {
        auto parts = split(s, '/');
        if (parts.size() < 2) throw ParseOptionError("Unexpected format in " + s);
        FadeMode mode;
        if (parts[0] == "repeat") mode = FadeMode::Repeat;
        else if (parts[0] == "clamp") mode = FadeMode::Clamp;
        else throw ParseEnumError("Not a fade mode: " + parts[0] + ". Must be one of (repeat|clamp)");
        bool absolute;
        if (parts[1] == "absolute" || parts[1] == "abs") absolute = true;
        else if (parts[1] == "relative" || parts[1] == "rel") absolute = false;
        else throw ParseOptionError("Not a frame fade spec: " + s);
        std::vector<std::pair<int, float>> frames;
        for (size_t i = 2; i < parts.size(); i++) {
            auto fv = split(parts[i], ',');
            if (fv.size() != 2) throw ParseOptionError("Expected (int,float) per frame for fade. Got: " + s);
            frames.emplace_back(parse_i32(fv[0], "Expected (int,float) per frame for fade. Got: " + s), parse_f32(fv[1], "Expected (int,float) per frame for fade. Got: " + s));
        }
        return Fade(mode, absolute, frames);
    }
    bool absolute() const { return absolute_; }
    chb_fade to_c() const {
        chb_fade f{};
        f.is_none = is_none_ ? 1 : 0;
        f.mode = (uint8_t)mode_;
        f.absolute = absolute_ ? 1 : 0;
        f.offset = offset_;
        f.n_values = (int32_t)values_.size();
        f.values = values_.empty() ? nullptr : values_.data();
        return f;
    }

private:
    Fade() : is_none_(true), mode_(FadeMode::Clamp), absolute_(true) {}
    bool is_none_;
    FadeMode mode_;
    bool absolute_;
    int32_t offset_ = 0;
    std::vector<float> values_;
};

// src/flist.rs:9-79
struct FrameRange {
    std::optional<int> start, end;
    unsigned step = 1;
    static FrameRange empty() { return FrameRange{}; }
    std::optional<int> range() const { return (start && end) ? std::optional<int>(*end - *start) : std::nullopt; }
    static FrameRange from_str(const std::string& s) {
        auto parts = split(s, '/');
        if (parts.size() != 3) throw ParseOptionError("Option --frames expects 3 elements: start/end/step, " + std::to_string(parts.size()) + " were suppied");
        FrameRange r;
        std::optional<int> v[3];
        for (int i = 0; i < 3; i++)
            if (parts[i] != ".") v[i] = parse_i32(parts[i], "Can't parse element " + std::to_string(i) + " in option --frames (start/end/step), got '" + parts[i] + "'.");
        r.start = v[0];
        r.end = v[1];
        r.step = (unsigned)v[2].value_or(1);
        return r;
    }
};

class Context {
public:
    explicit Context(const std::vector<int>& devices = {}) { check(chb_ctx_create(devices.empty() ? nullptr : devices.data(), (int)devices.size(), &h_)); }
    ~Context() { chb_ctx_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    chb_ctx* raw() const { return h_; }
    size_t free_bytes(int dev_slot = 0) const {
        size_t f = 0, t = 0;
        check(chb_ctx_mem_info(h_, dev_slot, &f, &t));
        return f;
    }

private:
    chb_ctx* h_ = nullptr;
};

// The HBM-resident frame stack: what the temp slice files are to the reference.
class GpuStack {
public:
    GpuStack(const Context& ctx, int width, int height, int channels, int n_frames) : w_(width), h_(height), c_(channels), n_(n_frames) {
        check(chb_stack_create(ctx.raw(), width, height, channels, n_frames, &s_));
    }
    ~GpuStack() { chb_stack_destroy(s_); }
    GpuStack(const GpuStack&) = delete;
    GpuStack& operator=(const GpuStack&) = delete;
    void upload(int frame_idx, const uint8_t* pixels, size_t row_pitch, int crop_x = 0, int crop_y = 0) { check(chb_stack_upload(s_, frame_idx, pixels, row_pitch, crop_x, crop_y)); }
    void sync() { check(chb_stack_sync(s_)); }
    chb_stack* raw() const { return s_; }
    int width() const { return w_; }
    int height() const { return h_; }
    int channels() const { return c_; }
    int frames() const { return n_; }
    size_t image_bytes() const { return (size_t)w_ * h_ * c_; }

private:
    chb_stack* s_ = nullptr;
    int w_, h_, c_, n_;
};

// src/chrono.rs:32-206
class OutlierProcessor {
public:
    OutlierProcessor(Threshold threshold, BackgroundMode bg_mode, OutlierSelectionMode outlier_mode, const float (&weights)[4], Fade fade,
                     std::optional<size_t> sample_count, uint64_t seed = 0)
        : threshold_(threshold), bg_(bg_mode), om_(outlier_mode), fade_(std::move(fade)), sample_(sample_count), seed_(seed) {
        for (int i = 0; i < 4; i++) w_[i] = weights[i];
    }
    // -> (buffer, is_outlier); image_indices = the window of a video frame (src/chrono.rs:102-139)
    std::pair<std::vector<uint8_t>, std::vector<uint8_t>> process(const GpuStack& stack, const std::vector<int32_t>* image_indices = nullptr) {
        chb_outlier_params p = params();
        std::vector<uint8_t> buffer(stack.image_bytes()), is_outlier(stack.image_bytes());
        uint64_t warnings = 0;
        check(chb_outlier(stack.raw(), &p, image_indices ? image_indices->data() : nullptr, image_indices ? (int)image_indices->size() : 0,
                          buffer.data(), is_outlier.data(), &warnings));
        warnings_ = warnings;
        return {std::move(buffer), std::move(is_outlier)};
    }
    uint64_t warnings() const { return warnings_; }  // "pixels seem to consist of only outliers", src/chrono.rs:198-203
    // global index of the stack's first pixel when the stack is one row band of a larger image (keeps the per-pixel RNG of
    // --background random / --sample aligned with the whole-image result)
    void set_pixel_offset(uint64_t offset) { pixel_offset_ = offset; }

    // chrono-video: n_windows windows of window_len consecutive frames, window i starting at frame first_start + i -- the
    // per-frame loop of create_video (src/main.rs:254-331) in one call (chb_outlier_video). Planes are [n_windows][H*W*C].
    static constexpr int kMaxVideoWindow = 64;
    bool slidable(int window_len) const { return window_len >= 1 && window_len <= kMaxVideoWindow && (!sample_ || (int)*sample_ >= window_len); }
    void process_video_run(const GpuStack& stack, int first_start, int window_len, int n_windows, std::vector<uint8_t>& buffers,
                           std::vector<uint8_t>& is_outlier, std::vector<uint64_t>& warnings) {
        chb_outlier_params p = params();
        buffers.resize(stack.image_bytes() * (size_t)n_windows);
        is_outlier.resize(stack.image_bytes() * (size_t)n_windows);
        warnings.assign((size_t)n_windows, 0);
        check(chb_outlier_video(stack.raw(), &p, first_start, window_len, n_windows, buffers.data(), is_outlier.data(), warnings.data()));
    }

private:
    chb_outlier_params params() const {
        chb_outlier_params p{};
        p.thr_absolute = threshold_.absolute() ? 1 : 0;
        p.background = (uint8_t)bg_;
        p.outlier = (uint8_t)om_;
        p.thr_min = threshold_.min(); p.thr_max = threshold_.max(); p.thr_scale = threshold_.scale();
        for (int i = 0; i < 4; i++) p.weights[i] = w_[i];
        p.fade = fade_.to_c();
        p.sample_count = sample_ ? (int32_t)*sample_ : -1;
        p.seed = seed_;
        p.pixel_offset = pixel_offset_;
        return p;
    }
    Threshold threshold_;
    BackgroundMode bg_;
    OutlierSelectionMode om_;
    float w_[4];
    Fade fade_;
    std::optional<size_t> sample_;
    uint64_t seed_, warnings_ = 0, pixel_offset_ = 0;
};

// src/simple.rs:11-168
class SimpleProcessor {
public:
    SimpleProcessor(const float (&weights)[4], Fade fade, bool darker) : fade_(std::move(fade)), darker_(darker) {
        for (int i = 0; i < 4; i++) w_[i] = weights[i];
    }
    std::vector<uint8_t> process(const GpuStack& stack, const std::vector<int32_t>* image_indices = nullptr) {
        chb_simple_params p{};
        p.darker = darker_ ? 1 : 0;
        for (int i = 0; i < 4; i++) p.weights[i] = w_[i];
        p.fade = fade_.to_c();
        std::vector<uint8_t> buffer(stack.image_bytes());
        check(chb_simple(stack.raw(), &p, image_indices ? image_indices->data() : nullptr, image_indices ? (int)image_indices->size() : 0, buffer.data()));
        return buffer;
    }

private:
    float w_[4];
    Fade fade_;
    bool darker_;
};

// src/shake.rs:44-119: `--shake <anchor-radius>/<search-radius>` and `--shake-anchors x/y`
struct ShakeParams {
    uint32_t anchor_radius = 0, search_radius = 0;
    static ShakeParams from_str(const std::string& s) {
        auto parts = split(s, '/');
        if (parts.size() != 2) throw ParseOptionError("Unexpected format in shake parameters, expected <rad>/<search-rad>: " + s);
        ShakeParams p;
        int a = parse_i32(parts[0], "Unexpected format in shake parameter: " + s), b = parse_i32(parts[1], "Unexpected format in shake parameter: " + s);
        if (a < 0 || b < 0) throw ParseOptionError("Unexpected format in shake parameter: " + s);
        p.anchor_radius = (uint32_t)a; p.search_radius = (uint32_t)b;
        return p;
    }
};
struct ShakeAnchor {
    int32_t x = 0, y = 0;
    static ShakeAnchor from_str(const std::string& s) {
        auto parts = split(s, '/');
        if (parts.size() != 2) throw ParseOptionError("Unexpected format in shake anchor, expected x/y: " + s);
        ShakeAnchor a;
        a.x = parse_i32(parts[0], "Unexpected format in shake anchor, expected x/y: " + s);
        a.y = parse_i32(parts[1], "Unexpected format in shake anchor, expected x/y: " + s);
        return a;
    }
};

// ShakeAnalyzer::analyze (src/shake.rs:190-305): the first frame supplies the anchor windows; offset() is one iteration of
// the reference's par_iter over the remaining frames (sums of squared differences on the GPU, first minimum).
class ShakeAnalyzer {
public:
    ShakeAnalyzer(const Context& ctx, int width, int height, int channels, const std::vector<ShakeAnchor>& anchors, const ShakeParams& params,
                  const uint8_t* first_frame, size_t row_pitch) {
        std::vector<int32_t> xy;
        for (const auto& a : anchors) { xy.push_back(a.x); xy.push_back(a.y); }
        check(chb_shake_create(ctx.raw(), width, height, channels, xy.data(), (int)anchors.size(), (int)params.anchor_radius, (int)params.search_radius,
                               first_frame, row_pitch, &h_));
    }
    ~ShakeAnalyzer() { chb_shake_destroy(h_); }
    ShakeAnalyzer(const ShakeAnalyzer&) = delete;
    ShakeAnalyzer& operator=(const ShakeAnalyzer&) = delete;
    std::pair<int32_t, int32_t> offset(const uint8_t* frame, size_t row_pitch) const {
        int32_t dx = 0, dy = 0;
        check(chb_shake_offset(h_, frame, row_pitch, &dx, &dy, nullptr));
        return {dx, dy};
    }

private:
    chb_shake* h_ = nullptr;
};

}  // namespace chrono_b200
