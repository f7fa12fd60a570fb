// chrono_b200_cli -- CLI-compatible driver for the compositing path (mirrors src/cli.rs:19-243 and src/main.rs:21-571
// for the hot-path flags). Frames: JPEG (decoded on the GPU by nvJPEG: the compressed bytes cross PCIe), PNG (RGB8 / RGBA8,
// zlib) or binary PPM; output by extension like save_image (src/main.rs:520-571): jpg/jpeg at --quality, png, tif/tiff, bmp,
// ppm. --shake / --shake-anchors run the camera-shake analysis on the GPU (src/main.rs:61-77) and the Crop origins are applied
// while the frames are uploaded. Flags that only steer the reference's temp files (--slice, --compression, --temp-dir) or its
// thread pools are accepted and ignored.
#include <glob.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>

#include "../include/chrono_b200.hpp"
#include "../include/chrono_b200_imageio.hpp"

using namespace chrono_b200;

// Cli::from_str quote handling (src/cli.rs:245-266)
static std::vector<std::string> split_option_string(const std::string& str) {
    std::vector<std::string> args;
    auto parts = split(str, '"');
    for (size_t i = 0; i < parts.size(); i++) {
        if (i % 2 == 0) {
            std::istringstream is(parts[i]);
            std::string tok;
            while (is >> tok) args.push_back(tok);
        } else {
            std::string t = parts[i];
            t.erase(0, t.find_first_not_of(" \t"));
            t.erase(t.find_last_not_of(" \t") + 1);
            args.push_back(t);
        }
    }
    return args;
}

static std::pair<std::string, std::string> name_and_extension(const std::string& path) {  // src/main.rs:431-451
    size_t slash = path.find_last_of('/');
    std::string file = slash == std::string::npos ? path : path.substr(slash + 1);
    size_t dot = file.find_last_of('.');
    if (dot == std::string::npos || dot == 0) throw std::runtime_error("Unexpected format in " + path);
    return {file.substr(0, dot), file.substr(dot + 1)};
}
static std::string parent_of(const std::string& path) {
    size_t slash = path.find_last_of('/');
    return slash == std::string::npos ? std::string(".") : path.substr(0, slash);
}

int main(int argc, char** argv) {
    try {
        std::vector<std::string> args(argv + 1, argv + argc);
        if (args.size() == 1 && args[0][0] != '-') {  // option file (src/main.rs:33-43)
            std::ifstream f(args[0]);
            if (!f) throw std::runtime_error("Something went wrong reading the options file " + args[0]);
            std::stringstream ss;
            ss << f.rdbuf();
            std::string content = ss.str();
            std::replace(content.begin(), content.end(), '\r', ' ');
            std::replace(content.begin(), content.end(), '\n', ' ');
            args = split_option_string(content);
        }
        std::map<std::string, std::string> opt;
        std::vector<float> weights;
        std::vector<std::string> anchor_args;
        const std::map<std::string, std::string> shorts = {{"-p", "--pattern"}, {"-f", "--frames"}, {"-o", "--output"}, {"-m", "--mode"}, {"-t", "--threshold"},
                                                           {"-b", "--background"}, {"-l", "--outlier"}, {"-c", "--compression"}, {"-q", "--quality"}, {"-s", "--slice"}};
        for (size_t i = 0; i < args.size(); i++) {
            std::string a = args[i];
            if (shorts.count(a)) a = shorts.at(a);
            if (a == "--debug" || a == "-d" || a == "--wait" || a == "-w") continue;
            if (a == "--weights") {
                for (int k = 0; k < 4; k++) {
                    if (++i >= args.size()) throw ParseOptionError("--weights requires 4 values");
                    weights.push_back(parse_f32(args[i], "Can't parse weight " + args[i]));
                }
                continue;
            }
            if (a == "--shake-anchors") {  // one or more x/y values (src/cli.rs:125-127)
                while (i + 1 < args.size() && args[i + 1].rfind("--", 0) != 0 && !(args[i + 1].size() == 2 && args[i + 1][0] == '-')) anchor_args.push_back(args[++i]);
                if (anchor_args.empty()) throw ParseOptionError("The argument '--shake-anchors' requires a value");
                opt[a] = "given";
                continue;
            }
            if (a.rfind("--", 0) != 0) throw ParseOptionError("Found argument '" + a + "' which wasn't expected");
            if (i + 1 >= args.size()) throw ParseOptionError("The argument '" + a + "' requires a value");
            opt[a] = args[++i];
        }
        if (!opt.count("--pattern") || !opt.count("--output")) throw ParseOptionError("The following required arguments were not provided: --pattern --output");
        // defaults: src/cli.rs:192-216
        SelectionMode mode = opt.count("--mode") ? parse_selection_mode(opt["--mode"]) : SelectionMode::Outlier;
        Threshold threshold = opt.count("--threshold") ? Threshold::from_str(opt["--threshold"]) : Threshold::abs(0.05f, 0.2f);
        BackgroundMode background = opt.count("--background") ? parse_background_mode(opt["--background"]) : BackgroundMode::Random;
        OutlierSelectionMode outlier = opt.count("--outlier") ? parse_outlier_mode(opt["--outlier"]) : OutlierSelectionMode::Extreme;
        Fade fade = opt.count("--fade") ? Fade::from_str(opt["--fade"]) : Fade::none();
        float w[4] = {1.0f, 1.0f, 1.0f, 1.0f};
        for (size_t i = 0; i < weights.size() && i < 4; i++) w[i] = weights[i];
        std::optional<size_t> sample;
        if (opt.count("--sample")) sample = (size_t)parse_i32(opt["--sample"], "Can't parse --sample");
        std::optional<FrameRange> frames, video_in, video_out;
        if (opt.count("--frames")) frames = FrameRange::from_str(opt["--frames"]);
        if (opt.count("--video-in")) video_in = FrameRange::from_str(opt["--video-in"]);
        if (opt.count("--video-out")) video_out = FrameRange::from_str(opt["--video-out"]);
        if (opt.count("--shake") != opt.count("--shake-anchors")) throw ParseOptionError("Provide both options or none: `--shake` and `--shake-anchors`");
        int quality = opt.count("--quality") ? parse_i32(opt["--quality"], "Can't parse --quality") : 95;  // src/cli.rs:197-200
        if (quality < 1 || quality > 100) throw ParseOptionError("--quality must be in 1..100");
        if (mode != SelectionMode::Outlier) {  // src/cli.rs:141-167, :233-239
            std::vector<std::string> unused;
            for (const char* k : {"--output-blend", "--threshold", "--outlier", "--background", "--temp-dir", "--sample", "--slice", "--compression"})
                if (opt.count(k)) unused.push_back(k);
            if (!unused.empty()) {
                std::cout << "WARNING! The following options are not used, as they are required only for `--mode outlier`:\n";
                for (auto& u : unused) std::cout << u << "\n";
                std::cout << "\n";
            }
        }
        // FileLister::files_vec (src/flist.rs:121-143): glob, then take(end).skip(start), then every step-th
        glob_t g;
        std::vector<std::string> files;
        if (glob(opt["--pattern"].c_str(), 0, nullptr, &g) == 0) {
            for (size_t i = 0; i < g.gl_pathc; i++) files.push_back(g.gl_pathv[i]);
        }
        globfree(&g);
        std::sort(files.begin(), files.end());
        if (frames) {
            size_t end = frames->end ? (size_t)std::max(0, *frames->end) : files.size();
            size_t start = frames->start ? (size_t)std::max(0, *frames->start) : 0;
            std::vector<std::string> sel;
            for (size_t i = start, k = 0; i < std::min(end, files.size()); i++, k++)
                if (k % frames->step == 0) sel.push_back(files[i]);
            files.swap(sel);
        }
        if (files.empty()) throw std::runtime_error("Unable to process search pattern " + opt["--pattern"]);

        // ---- camera-shake analysis (src/main.rs:61-84: ShakeAnalyzer::analyze -> Crop::create), then the upload that replaces
        // to_time_slices (src/main.rs:574-602): every frame lands in the HBM-resident stack at its Crop origin
        Context ctx;
        Image first = read_image(files[0], &ctx);
        const int fw = first.w, fh = first.h, ch = first.c;
        std::vector<int32_t> origins(2 * files.size(), 0);
        int cw = fw, chh = fh;
        if (opt.count("--shake")) {
            const ShakeParams sp = ShakeParams::from_str(opt["--shake"]);
            std::vector<ShakeAnchor> anchors;
            for (const auto& s : anchor_args) anchors.push_back(ShakeAnchor::from_str(s));
            std::cout << "Analyzing camera shake in " << files.size() << " images\n";
            ShakeAnalyzer an(ctx, fw, fh, ch, anchors, sp, first.px.data(), (size_t)fw * ch);
            std::vector<int32_t> offs(2 * files.size(), 0);  // the first frame is the reference: offset (0, 0) (src/shake.rs:284-286)
            for (size_t i = 1; i < files.size(); i++) {
                Image im = read_image(files[i], &ctx);
                if (im.w != fw || im.h != fh || im.c != ch) throw std::runtime_error("Image layout does not fit!");
                auto o = an.offset(im.px.data(), (size_t)im.w * im.c);
                offs[2 * i] = o.first; offs[2 * i + 1] = o.second;
            }
            int32_t w2 = 0, h2 = 0;
            if (chb_crop_create(offs.data(), (int)files.size(), fw, fh, origins.data(), &w2, &h2)) {
                std::cout << "Camera shake detected. Images will be corrected.\n";
                cw = w2; chh = h2;
                if (cw < 1 || chh < 1) throw std::runtime_error("Camera shake larger than the image");
            } else {
                std::cout << "No camera shake detected. Images will not be corrected.\n";
                std::fill(origins.begin(), origins.end(), 0);
            }
        }
        // Fills a stack of `rows` rows with rows [row0, row0 + rows) of every (cropped) frame
        auto fill_stack = [&](GpuStack& stack, int row0) {
            for (size_t i = 0; i < files.size(); i++) {
                const int ox = origins[2 * i], oy = origins[2 * i + 1] + row0;
                if (is_jpeg_path(files[i]) && ch == 3) {  // compressed bytes to the device, decoded there (chb_stack_upload_jpeg checks the layout)
                    const std::vector<uint8_t> bytes = read_file(files[i]);
                    check(chb_stack_upload_jpeg(stack.raw(), (int)i, bytes.data(), bytes.size(), ox, oy));
                } else {
                    Image im = i == 0 ? first : read_image(files[i], &ctx);
                    if (im.w != fw || im.h != fh || im.c != ch) throw std::runtime_error("Image layout does not fit!");  // src/simple.rs:62-66
                    stack.upload((int)i, im.px.data(), (size_t)im.w * im.c, ox, oy);
                }
            }
            stack.sync();
        };
        auto write_out = [&](const std::string& path, const uint8_t* px) { save_image(px, cw, chh, ch, path, quality, &ctx); };
        std::optional<std::string> out_blend;
        if (opt.count("--output-blend")) out_blend = opt["--output-blend"];
        const bool is_video = video_in || video_out;

        // Out of core, like the reference's time slices (SliceLength::bytes, src/slicer.rs:19-41): when the series does not fit the
        // device, the image is composited in row bands, each band a stack of its own (the frames are read once per band).
        // CHRONO_B200_BAND_ROWS forces a band height (tests).
        const int n_groups = ((int)files.size() + 15) / 16;
        const size_t bytes_per_row = (size_t)cw * ((size_t)ch * n_groups * 16 + 34 * (size_t)ch + 80);  // stack + staging + planes + queues
        int band_rows = chh;
        {
            const size_t budget = (size_t)(0.85 * (double)ctx.free_bytes());
            if (bytes_per_row * (size_t)chh > budget) band_rows = (int)std::max<size_t>(1, budget / bytes_per_row);
            if (const char* e = getenv("CHRONO_B200_BAND_ROWS")) band_rows = std::max(1, std::min(chh, atoi(e)));
        }
        if (band_rows < chh) {
            if (is_video) throw std::runtime_error("The frame series does not fit the device; video output needs the whole clip resident");
            std::cout << "Processing in row bands of " << band_rows << " rows\n";
            const size_t row_bytes = (size_t)cw * ch;
            std::vector<uint8_t> image(row_bytes * chh), blend(mode == SelectionMode::Outlier ? row_bytes * chh : 0);
            uint64_t warnings = 0;
            for (int row0 = 0; row0 < chh; row0 += band_rows) {
                const int rows = std::min(band_rows, chh - row0);
                GpuStack band(ctx, cw, rows, ch, (int)files.size());
                fill_stack(band, row0);
                if (mode == SelectionMode::Outlier) {
                    OutlierProcessor proc(threshold, background, outlier, w, fade, sample, /*seed=*/0x9E3779B97F4A7C15ULL);
                    proc.set_pixel_offset((uint64_t)row0 * cw);
                    auto res = proc.process(band, nullptr);
                    warnings += proc.warnings();
                    std::memcpy(&image[row_bytes * row0], res.first.data(), row_bytes * rows);
                    std::memcpy(&blend[row_bytes * row0], res.second.data(), row_bytes * rows);
                } else {
                    SimpleProcessor proc(w, fade, mode == SelectionMode::Darker);
                    auto res = proc.process(band, nullptr);
                    std::memcpy(&image[row_bytes * row0], res.data(), row_bytes * rows);
                }
            }
            if (warnings > 0) std::cout << "Warning: " << warnings << " pixels seem to consist of only outliers\n";
            write_out(opt["--output"], image.data());
            if (out_blend && mode == SelectionMode::Outlier) write_out(*out_blend, blend.data());
            return 0;
        }

        GpuStack stack(ctx, cw, chh, ch, (int)files.size());
        fill_stack(stack, 0);
        auto run_frame = [&](const std::vector<int32_t>* indices, const std::string& out, const std::optional<std::string>& out_blend) {
            if (mode == SelectionMode::Outlier) {
                OutlierProcessor proc(threshold, background, outlier, w, fade, sample, /*seed=*/0x9E3779B97F4A7C15ULL);
                auto res = proc.process(stack, indices);
                if (proc.warnings() > 0) std::cout << "Warning: " << proc.warnings() << " pixels seem to consist of only outliers\n";
                write_out(out, res.first.data());
                if (out_blend) write_out(*out_blend, res.second.data());
            } else {
                SimpleProcessor proc(w, fade, mode == SelectionMode::Darker);
                write_out(out, proc.process(stack, indices).data());
            }
        };
        if (is_video) {  // create_video / create_video_simple (src/main.rs:214-429)
            FrameRange vin = video_in.value_or(FrameRange::empty()), vout = video_out.value_or(FrameRange::empty());
            const int cap = 4 * (int)files.size() + 16;
            std::vector<int32_t> ws(cap), we(cap), num(cap);
            int n = chb_video_windows((int)files.size(), vin.start.has_value(), vin.start.value_or(0), vin.end.has_value(), vin.end.value_or(0), (int)vin.step,
                                      vout.start.has_value(), vout.start.value_or(0), vout.end.has_value(), vout.end.value_or(0), (int)vout.step, ws.data(),
                                      we.data(), num.data(), cap);
            std::string dir = parent_of(opt["--output"]);
            auto frame_name = [&](const std::string& base, int number) {
                char nm[64];
                snprintf(nm, sizeof nm, "-%05d.", number);
                auto nb = name_and_extension(base);
                return dir + "/" + nb.first + nm + nb.second;
            };
            const int nwin = std::min(n, cap);
            for (int i = 0; i < nwin;) {
                const int len = (we[i] - ws[i] + (int)vin.step - 1) / (int)vin.step;
                if (len <= 0) { std::cout << "Skipping frame " << num[i] << "\n"; i++; continue; }
                // a run of equal-length windows sliding by one frame goes through the sliding-window kernel in one call
                int j = i + 1;
                OutlierProcessor vproc(threshold, background, outlier, w, fade, sample, /*seed=*/0x9E3779B97F4A7C15ULL);
                if (mode == SelectionMode::Outlier && vin.step == 1 && vproc.slidable(len))
                    while (j < nwin && we[j] - ws[j] == len && ws[j] == ws[j - 1] + 1) j++;
                if (j - i > 1) {
                    std::vector<uint8_t> bufs, masks;
                    std::vector<uint64_t> warns;
                    vproc.process_video_run(stack, ws[i], len, j - i, bufs, masks, warns);
                    const size_t fb = (size_t)cw * chh * ch;
                    for (int k = i; k < j; k++) {
                        std::cout << "Processing frame " << num[k] << " -> \n";
                        if (warns[k - i] > 0) std::cout << "Warning: " << warns[k - i] << " pixels seem to consist of only outliers\n";
                        write_out(frame_name(opt["--output"], num[k]), bufs.data() + (k - i) * fb);
                        if (out_blend) write_out(frame_name(*out_blend, num[k]), masks.data() + (k - i) * fb);
                    }
                    i = j;
                    continue;
                }
                std::vector<int32_t> idx;
                for (int f = ws[i]; f < we[i]; f += (int)vin.step) idx.push_back(f);
                std::optional<std::string> ob;
                if (out_blend) ob = frame_name(*out_blend, num[i]);
                std::cout << "Processing frame " << num[i] << " -> \n";
                run_frame(&idx, frame_name(opt["--output"], num[i]), ob);
                i++;
            }
        } else {
            run_frame(nullptr, opt["--output"], out_blend);
        }
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << "\n";
        return 1;
    }
}
