// EXR output / input, behaviour of Lumen's ImageUtils (reference: src/Framework/ImageUtils.cpp:8-89):
// RGBA fp32 in memory -> planar B, G, R channels stored as HALF through tinyexr (+ miniz, CMakeLists.txt:52).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define TINYEXR_USE_MINIZ 1
#define TINYEXR_IMPLEMENTATION
#include <miniz.h>
#include <tinyexr.h>

#include "lumen_scene.h"

namespace lmh {

bool save_exr(const float* rgba, int width, int height, const char* path, std::string* err_out) {
	EXRHeader header;
	InitEXRHeader(&header);
	EXRImage image;
	InitEXRImage(&image);
	image.num_channels = 3;
	const size_t n = (size_t)width * height;
	std::vector<float> planes[3];
	for (int c = 0; c < 3; c++) {
		planes[c].resize(n);
		for (size_t i = 0; i < n; i++) planes[c][i] = rgba[4 * i + c];
	}
	float* image_ptr[3] = {planes[2].data(), planes[1].data(), planes[0].data()};  // B, G, R
	image.images = (unsigned char**)image_ptr;
	image.width = width;
	image.height = height;
	header.num_channels = 3;
	header.channels = (EXRChannelInfo*)malloc(sizeof(EXRChannelInfo) * 3);
	const char* names[3] = {"B", "G", "R"};
	for (int c = 0; c < 3; c++) {
		strncpy(header.channels[c].name, names[c], 255);
		header.channels[c].name[1] = '\0';
	}
	header.pixel_types = (int*)malloc(sizeof(int) * 3);
	header.requested_pixel_types = (int*)malloc(sizeof(int) * 3);
	for (int c = 0; c < 3; c++) {
		header.pixel_types[c] = TINYEXR_PIXELTYPE_FLOAT;
		header.requested_pixel_types[c] = TINYEXR_PIXELTYPE_HALF;  // ImageUtils.cpp:72-76
	}
	const char* err = nullptr;
	const int ret = SaveEXRImageToFile(&image, &header, path, &err);
	if (ret != TINYEXR_SUCCESS && err_out) *err_out = err ? err : "SaveEXRImageToFile failed";
	if (err) FreeEXRErrorMessage(err);
	free(header.channels);
	free(header.pixel_types);
	free(header.requested_pixel_types);
	return ret == TINYEXR_SUCCESS;
}

// Same file, from channel planes already converted to HALF on the device (lmb_download_half_bgr): planes = B, G, R,
// width*height halves each. tinyexr copies HALF input to a HALF channel unchanged, so the bytes written equal save_exr's.
bool save_exr_half_bgr(const uint16_t* planes, int width, int height, const char* path, std::string* err_out) {
	EXRHeader header;
	InitEXRHeader(&header);
	EXRImage image;
	InitEXRImage(&image);
	image.num_channels = 3;
	const size_t n = (size_t)width * height;
	const uint16_t* image_ptr[3] = {planes, planes + n, planes + 2 * n};
	image.images = (unsigned char**)image_ptr;
	image.width = width;
	image.height = height;
	header.num_channels = 3;
	header.channels = (EXRChannelInfo*)malloc(sizeof(EXRChannelInfo) * 3);
	const char* names[3] = {"B", "G", "R"};
	for (int c = 0; c < 3; c++) {
		strncpy(header.channels[c].name, names[c], 255);
		header.channels[c].name[1] = '\0';
	}
	header.pixel_types = (int*)malloc(sizeof(int) * 3);
	header.requested_pixel_types = (int*)malloc(sizeof(int) * 3);
	for (int c = 0; c < 3; c++) header.pixel_types[c] = header.requested_pixel_types[c] = TINYEXR_PIXELTYPE_HALF;
	const char* err = nullptr;
	const int ret = SaveEXRImageToFile(&image, &header, path, &err);
	if (ret != TINYEXR_SUCCESS && err_out) *err_out = err ? err : "SaveEXRImageToFile failed";
	if (err) FreeEXRErrorMessage(err);
	free(header.channels);
	free(header.pixel_types);
	free(header.requested_pixel_types);
	return ret == TINYEXR_SUCCESS;
}

bool save_checkpoint(const char* path, const float* rgba, uint32_t width, uint32_t height, uint32_t frames, uint32_t path_length, std::string* err) {
	const std::string tmp = std::string(path) + ".tmp";  // write-then-rename: an interrupted save never clobbers the last good checkpoint
	FILE* f = fopen(tmp.c_str(), "wb");
	if (!f) {
		if (err) *err = "cannot open " + tmp;
		return false;
	}
	const uint32_t hdr[6] = {0x4B434D4Cu /* "LMCK" */, 1u, width, height, frames, path_length};
	const size_t n = (size_t)width * height * 4;
	const bool ok = fwrite(hdr, sizeof(hdr), 1, f) == 1 && fwrite(rgba, sizeof(float), n, f) == n;
	if (fclose(f) != 0 || !ok || rename(tmp.c_str(), path) != 0) {
		if (err) *err = std::string("write failed: ") + path;
		remove(tmp.c_str());
		return false;
	}
	return true;
}

bool load_checkpoint(const char* path, std::vector<float>& rgba, uint32_t& width, uint32_t& height, uint32_t& frames, uint32_t& path_length,
					 std::string* err) {
	FILE* f = fopen(path, "rb");
	if (!f) {
		if (err) *err = std::string("cannot open ") + path;
		return false;
	}
	uint32_t hdr[6] = {};
	bool ok = fread(hdr, sizeof(hdr), 1, f) == 1 && hdr[0] == 0x4B434D4Cu && hdr[1] == 1u && hdr[2] > 0 && hdr[3] > 0 && (uint64_t)hdr[2] * hdr[3] <= (1ull << 28);
	if (ok) {
		const size_t n = (size_t)hdr[2] * hdr[3] * 4;
		rgba.resize(n);
		ok = fread(rgba.data(), sizeof(float), n, f) == n && fgetc(f) == EOF;
	}
	fclose(f);
	if (!ok) {
		if (err) *err = std::string("not a lumen_b200 checkpoint (or truncated): ") + path;
		return false;
	}
	width = hdr[2], height = hdr[3], frames = hdr[4], path_length = hdr[5];
	return true;
}

bool load_exr(const char* path, std::vector<float>& rgba, int& width, int& height, std::string* err_out) {
	const char* err = nullptr;
	float* data = nullptr;
	const int ret = LoadEXR(&data, &width, &height, path, &err);
	if (ret != TINYEXR_SUCCESS) {
		if (err_out) *err_out = err ? err : "LoadEXR failed";
		if (err) FreeEXRErrorMessage(err);
		return false;
	}
	rgba.assign(data, data + (size_t)width * height * 4);
	free(data);
	return true;
}

}  // namespace lmh
