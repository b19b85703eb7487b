// The reference's --debug artefacts of the corner detector, SURVEY.md row F4 (host code; nothing here is on a hot path):
//   /tmp/mrgingham-scaled-processed-level%d.png              the pyramid-level image             find_chessboard_corners.cc:452-459
//   /tmp/mrgingham-chess-response[-refinement]-level%d.png   ChESS response, normalised to 0..255                          :513-523
//   /tmp/mrgingham-chess-response[-refinement]-level%d-positive.png   ... after negatives are zeroed                      :531-541
//   /tmp/mrgingham-1-corners.vnl / ...-refinement-level%d.vnl         self-plotting list of the corners found / refined    :294-315, :346-348, :391-392, :400-407
// The images are 8-bit greyscale PNGs with the pixel values OpenCV's cv::normalize + cv::imwrite give (the normalisation
// is convertTo's float multiply-add, rounded half to even; imwrite converts the 16-bit signed result to 8 bits by
// saturation); the files themselves are written with stored (uncompressed) deflate blocks, so their bytes differ from
// libpng's. The corner list holds the un-quantised doubles, printed with %f as the reference prints them.
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <sys/stat.h>
#include <vector>

#include "kernels.cuh"

namespace mrgb200
{
namespace
{
uint32_t crc_table[256];
bool crc_ready = false;
uint32_t crc32_of(const uint8_t* p, size_t n, uint32_t c = 0xFFFFFFFFu)
{
    if (!crc_ready)
    {
        for (uint32_t i = 0; i < 256; i++) { uint32_t v = i; for (int k = 0; k < 8; k++) v = (v & 1) ? 0xEDB88320u ^ (v >> 1) : v >> 1; crc_table[i] = v; }
        crc_ready = true;
    }
    for (size_t i = 0; i < n; i++) c = crc_table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return c;
}
void put32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
void chunk(std::vector<uint8_t>& out, const char type[4], const std::vector<uint8_t>& data)
{
    put32(out, (uint32_t)data.size());
    const size_t at = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), data.begin(), data.end());
    put32(out, crc32_of(out.data() + at, out.size() - at) ^ 0xFFFFFFFFu);
}
}   // namespace

// 8-bit greyscale PNG, deflate "stored" blocks
bool write_png_gray8(const char* path, const uint8_t* data, int w, int h, size_t pitch)
{
    std::vector<uint8_t> raw;
    raw.reserve((size_t)(w + 1) * h);
    for (int y = 0; y < h; y++) { raw.push_back(0); raw.insert(raw.end(), data + (size_t)y * pitch, data + (size_t)y * pitch + w); }
    std::vector<uint8_t> z;
    z.push_back(0x78); z.push_back(0x01);
    uint32_t a = 1, b = 0;
    for (size_t i = 0; i < raw.size(); i++) { a = (a + raw[i]) % 65521u; b = (b + a) % 65521u; }
    for (size_t at = 0; at < raw.size() || at == 0; )
    {
        const size_t n = raw.size() - at < 65535 ? raw.size() - at : 65535;
        z.push_back(at + n >= raw.size() ? 1 : 0);
        z.push_back(n & 0xFF); z.push_back(n >> 8); z.push_back(~n & 0xFF); z.push_back((~n >> 8) & 0xFF);
        z.insert(z.end(), raw.begin() + at, raw.begin() + at + n);
        at += n;
        if (n == 0) break;
    }
    put32(z, (b << 16) | a);
    std::vector<uint8_t> out = { 0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A };
    std::vector<uint8_t> ihdr;
    put32(ihdr, (uint32_t)w); put32(ihdr, (uint32_t)h);
    ihdr.push_back(8); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    chunk(out, "IHDR", ihdr); chunk(out, "IDAT", z); chunk(out, "IEND", std::vector<uint8_t>());
    FILE* fp = fopen(path, "wb");
    if (!fp) return false;
    const bool ok = fwrite(out.data(), 1, out.size(), fp) == out.size();
    fclose(fp);
    return ok;
}

// cv::normalize(response /* CV_16S */, out, 0, 255, NORM_MINMAX) followed by cv::imwrite's conversion to 8 bits
void normalize_response_u8(const int16_t* resp, size_t n, uint8_t* out)
{
    int mn = 32767, mx = -32768;
    for (size_t i = 0; i < n; i++) { if (resp[i] < mn) mn = resp[i]; if (resp[i] > mx) mx = resp[i]; }
    const double smin = mn, smax = mx;
    const double scale = 255.0 * (smax - smin > DBL_EPSILON ? 1. / (smax - smin) : 0);
    const double shift = 0.0 - smin * scale;
    const float a = (float)scale, b = (float)shift;
    for (size_t i = 0; i < n; i++)
    {
        long q = lrintf(fmaf((float)resp[i], a, b));
        out[i] = (uint8_t)(q < 0 ? 0 : q > 255 ? 255 : q);
    }
}

// the self-plotting corner list
bool write_corner_vnl(const char* path, const char* debug_image_filename, const double* xy, int n)
{
    FILE* fp = fopen(path, "w");
    if (!fp) return false;
    if (debug_image_filename) fprintf(fp, "#!/usr/bin/feedgnuplot --dom --with 'points pt 7 ps 2' --square --image %s\n", debug_image_filename);
    else                      fprintf(fp, "#!/usr/bin/feedgnuplot --dom --square --set 'yr [:] rev'\n");
    fprintf(fp, "# x y\n");
    for (int i = 0; i < n; i++) fprintf(fp, "%f %f\n", xy[2*i], xy[2*i + 1]);
    fclose(fp);
    chmod(path, S_IRUSR | S_IRGRP | S_IROTH | S_IWUSR | S_IWGRP | S_IXUSR | S_IXGRP | S_IXOTH);
    return true;
}

}
