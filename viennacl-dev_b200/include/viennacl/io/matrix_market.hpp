// viennacl/io/matrix_market.hpp -- MatrixMarket coordinate-format reader / writer (reference: io/matrix_market.hpp:101-420),
// the ingest step in front of the hot path.  Same entry points and conventions as the reference:
//   long read_matrix_market_file(std::vector<std::map<unsigned int, T>> & mat, file, index_base = 1)
//     returns the number of lines read (> 0) on success, 0 on a malformed file, EXIT_FAILURE when the file cannot be opened;
//     `symmetric` headers mirror the off-diagonal entries, `pattern` files get the value 1, dense `array` files are read
//     column by column; the matrix is resized to the header's dimensions.
//   void write_matrix_market_file(mat, file, index_base = 1)      general real coordinate format
// plus overloads that read straight into a device compressed_matrix (host parse, one upload).
#ifndef VIENNACL_B200_IO_MATRIX_MARKET_HPP
#define VIENNACL_B200_IO_MATRIX_MARKET_HPP

#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include "viennacl/compressed_matrix.hpp"

namespace viennacl
{
namespace io
{
namespace detail
{
  inline std::string lower(std::string s)
  {
    for (std::size_t i = 0; i < s.size(); ++i) s[i] = static_cast<char>(std::tolower(static_cast<unsigned char>(s[i])));
    return s;
  }
}

template<typename ScalarT>
long read_matrix_market_file(std::vector< std::map<unsigned int, ScalarT> > & mat, const char *file, long index_base = 1)
{
  std::ifstream reader(file);
  if (!reader)
  {
    std::cerr << "ViennaCL: Matrix Market Reader: Cannot open file " << file << std::endl;
    return EXIT_FAILURE;
  }
  std::string line;
  long linenum = 0, rows = 0, cols = 0, nnz = 0, valid_entries = 0;
  bool symmetric = false, pattern = false, dense = false, have_sizes = false;
  long dense_row = 0, dense_col = 0;

  while (std::getline(reader, line))
  {
    ++linenum;
    if (linenum == 1)
    {
      // %%MatrixMarket matrix coordinate real general
      std::istringstream hs(detail::lower(line));
      std::string banner, object, format, field, symmetry;
      hs >> banner >> object >> format >> field >> symmetry;
      if (banner != "%%matrixmarket") { std::cerr << "Error in file " << file << " at line 1: expected '%%MatrixMarket'" << std::endl; return 0; }
      if (object != "matrix") { std::cerr << "Error in file " << file << ": expected 'matrix' in the header" << std::endl; return 0; }
      if (format == "array") dense = true;
      else if (format != "coordinate") { std::cerr << "Error in file " << file << ": format must be 'coordinate' or 'array'" << std::endl; return 0; }
      if (field == "pattern") pattern = true;
      else if (field == "complex") { std::cerr << "Error in file " << file << ": complex matrices are not supported" << std::endl; return 0; }
      else if (field != "real" && field != "integer" && field != "double") { std::cerr << "Error in file " << file << ": unknown field type '" << field << "'" << std::endl; return 0; }
      if (symmetry == "symmetric") symmetric = true;
      else if (symmetry != "general") { std::cerr << "Error in file " << file << ": symmetry must be 'general' or 'symmetric'" << std::endl; return 0; }
      continue;
    }
    std::size_t first = line.find_first_not_of(" \t\r");
    if (first == std::string::npos || line[first] == '%') continue;          // blank line or comment
    std::istringstream ls(line);
    if (!have_sizes)
    {
      if (dense) { if (!(ls >> rows >> cols)) { std::cerr << "Error in file " << file << " at line " << linenum << ": bad size line" << std::endl; return 0; } nnz = rows * cols; }
      else if (!(ls >> rows >> cols >> nnz)) { std::cerr << "Error in file " << file << " at line " << linenum << ": bad size line" << std::endl; return 0; }
      if (rows < 0 || cols < 0 || nnz < 0) { std::cerr << "Error in file " << file << ": negative dimensions" << std::endl; return 0; }
      mat.assign(static_cast<std::size_t>(rows), std::map<unsigned int, ScalarT>());
      have_sizes = true;
      continue;
    }
    if (dense)
    {
      ScalarT value;
      if (!(ls >> value)) { std::cerr << "Error in file " << file << " at line " << linenum << ": bad value" << std::endl; return 0; }
      if (dense_col >= cols) { std::cerr << "Error in file " << file << " at line " << linenum << ": too many entries" << std::endl; return 0; }
      if (value != ScalarT(0)) mat[static_cast<std::size_t>(dense_row)][static_cast<unsigned int>(dense_col)] = value;
      if (++dense_row == rows) { dense_row = 0; ++dense_col; }
      ++valid_entries;
      continue;
    }
    long r = 0, c = 0;
    ScalarT value = ScalarT(1);
    if (!(ls >> r >> c)) { std::cerr << "Error in file " << file << " at line " << linenum << ": expected row and column index" << std::endl; return 0; }
    if (!pattern && !(ls >> value)) { std::cerr << "Error in file " << file << " at line " << linenum << ": expected a value" << std::endl; return 0; }
    r -= index_base; c -= index_base;
    if (r < 0 || r >= rows || c < 0 || c >= cols)
    { std::cerr << "Error in file " << file << " at line " << linenum << ": index out of bounds (" << r + index_base << ", " << c + index_base << ")" << std::endl; return 0; }
    if (valid_entries >= nnz) { std::cerr << "Error in file " << file << " at line " << linenum << ": more entries than announced" << std::endl; return 0; }
    mat[static_cast<std::size_t>(r)][static_cast<unsigned int>(c)] = value;
    if (symmetric && r != c && c < rows && r < cols) mat[static_cast<std::size_t>(c)][static_cast<unsigned int>(r)] = value;
    ++valid_entries;
  }
  if (!have_sizes) { std::cerr << "Error in file " << file << ": no size line" << std::endl; return 0; }
  if (valid_entries != nnz) { std::cerr << "Error in file " << file << ": " << valid_entries << " entries read, " << nnz << " announced" << std::endl; return 0; }
  return linenum;
}

template<typename ScalarT>
long read_matrix_market_file(std::vector< std::map<unsigned int, ScalarT> > & mat, const std::string & file, long index_base = 1)
{ return read_matrix_market_file(mat, file.c_str(), index_base); }

/** @brief Reads straight into a device compressed_matrix: host parse, one upload; `cols` follows the file header */
template<typename NumericT, unsigned int AlignmentV>
long read_matrix_market_file(viennacl::compressed_matrix<NumericT, AlignmentV> & mat, const char *file, long index_base = 1)
{
  std::vector< std::map<unsigned int, NumericT> > host;
  // the header's column count is needed for the device matrix: re-parse it cheaply after a successful read
  const long lines = read_matrix_market_file(host, file, index_base);
  if (lines <= 0 || (lines == EXIT_FAILURE && host.empty())) return lines;
  long rows = 0, cols = 0;
  {
    std::ifstream reader(file);
    std::string line;
    std::getline(reader, line);
    while (std::getline(reader, line))
    {
      std::size_t first = line.find_first_not_of(" \t\r");
      if (first == std::string::npos || line[first] == '%') continue;
      std::istringstream ls(line);
      ls >> rows >> cols;
      break;
    }
  }
  vcl_size_t nnz = 0;
  for (std::size_t i = 0; i < host.size(); ++i) nnz += host[i].size();
  if (host.empty()) return lines;
  std::vector<unsigned int> rp(host.size() + 1), ci(nnz ? nnz : 1);
  std::vector<NumericT> va(nnz ? nnz : 1);
  vcl_size_t k = 0;
  for (std::size_t i = 0; i < host.size(); ++i)
  {
    rp[i] = static_cast<unsigned int>(k);
    for (typename std::map<unsigned int, NumericT>::const_iterator it = host[i].begin(); it != host[i].end(); ++it, ++k) { ci[k] = it->first; va[k] = it->second; }
  }
  rp[host.size()] = static_cast<unsigned int>(k);
  mat.set(&rp[0], &ci[0], &va[0], host.size(), static_cast<vcl_size_t>(cols), nnz);
  return lines;
}

template<typename NumericT, unsigned int AlignmentV>
long read_matrix_market_file(viennacl::compressed_matrix<NumericT, AlignmentV> & mat, const std::string & file, long index_base = 1)
{ return read_matrix_market_file(mat, file.c_str(), index_base); }

template<typename ScalarT>
void write_matrix_market_file(std::vector< std::map<unsigned int, ScalarT> > const & mat, const char *file, long index_base = 1)
{
  std::ofstream writer(file);
  vcl_size_t nnz = 0, cols = 0;
  for (std::size_t i = 0; i < mat.size(); ++i)
  {
    nnz += mat[i].size();
    if (!mat[i].empty()) cols = std::max<vcl_size_t>(cols, mat[i].rbegin()->first + 1);
  }
  writer << "%%MatrixMarket matrix coordinate real general" << std::endl;
  writer << mat.size() << " " << cols << " " << nnz << std::endl;
  writer.precision(17);
  for (std::size_t i = 0; i < mat.size(); ++i)
    for (typename std::map<unsigned int, ScalarT>::const_iterator it = mat[i].begin(); it != mat[i].end(); ++it)
      writer << (long(i) + index_base) << " " << (long(it->first) + index_base) << " " << it->second << std::endl;
}

template<typename ScalarT>
void write_matrix_market_file(std::vector< std::map<unsigned int, ScalarT> > const & mat, const std::string & file, long index_base = 1)
{ write_matrix_market_file(mat, file.c_str(), index_base); }

} // namespace io
} // namespace viennacl
#endif
