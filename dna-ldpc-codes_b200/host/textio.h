// textio.h - the text formats around the decoder: codeword / LLR readers, dec_*.txt writer.
#pragma once
#include <string>
#include <vector>

namespace dnaldpc {

// N integers 0/1 separated by whitespace (DNA_main.cpp:1322-1329, fscanf "%d").
bool read_codeword_txt(const std::string &path, int N, std::vector<signed char> &out, std::string &err);
// N decimal doubles = LLR ln(p0/p1) (DNA_main.cpp:1338-1343, fscanf "%lf").
bool read_llr_txt(const std::string &path, int N, std::vector<double> &out, std::string &err);
// N x "%d " of the decoded bits, no trailing newline (DNA_main.cpp:916-927).
bool write_dec_txt(const std::string &path, const unsigned char *bits, int N, std::string &err);

}  // namespace dnaldpc
