// 8-bit image input for the XML scene loader (PNG: image_io.cpp, JPEG: jpeg_read.cpp, Radiance HDR: hdr_read.cpp)
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace cptb {

struct Image8 {
    int width = 0, height = 0, channels = 0; // channels as stored in the file (1 grey, 2 grey+a, 3 rgb, 4 rgba)
    std::vector<uint8_t> data;
};

Image8 decode_png(const std::string &path); // throws std::runtime_error
// 8-bit channel conversion with the rules of the decoder the reference links (stb_image's convert_format)
std::vector<uint8_t> convert_channels(const Image8 &img, int want_channels);
// baseline / extended-sequential / progressive JPEG, decoded straight to `want_channels` with the rules of the decoder the reference
// links (1 / 2 channels from a colour file = its luma plane); jpeg_read.cpp
std::vector<uint8_t> decode_jpeg(const std::vector<uint8_t> &file, const std::string &path, int want_channels, int &width, int &height);
// Radiance RGBE (.hdr), tone-mapped to 8 bits like the reference's decoder does for its 8-bit entry point; hdr_read.cpp
std::vector<uint8_t> decode_hdr(const std::vector<uint8_t> &file, const std::string &path, int want_channels, int &width, int &height);
std::vector<uint8_t> load_image8(const std::string &path, int want_channels, int &width, int &height);

} // namespace cptb
