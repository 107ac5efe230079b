// EXR output / input, behaviour of Lumen's ImageUtils (reference: src/Framework/ImageUtils.cpp:8-89):
// RGBA fp32 in memory -> planar B, G, R channels stored as HALF through tinyexr (+ miniz, CMakeLists.txt:52).
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
