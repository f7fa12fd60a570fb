// imageio_convert <in> <out> -- reads a frame with the driver's reader and writes it with the driver's save_image
// (include/chrono_b200_imageio.hpp). Test helper for the CPU suite: PNG / PPM in, PNG / TIFF / BMP / PPM out need no GPU.
#include <iostream>

#include "../include/chrono_b200_imageio.hpp"

int main(int argc, char** argv) {
    if (argc != 3) { std::cerr << "usage: imageio_convert <in> <out>\n"; return 2; }
    try {
        const chrono_b200::Image im = chrono_b200::read_image(argv[1]);
        chrono_b200::save_image(im.px.data(), im.w, im.h, im.c, argv[2], 95, nullptr);
        std::cout << im.w << " " << im.h << " " << im.c << "\n";
        return 0;
    } catch (const std::exception& e) {
        std::cerr << "error: " << e.what() << "\n";
        return 1;
    }
}
