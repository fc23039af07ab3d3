// viennacl/tools/matrix_generation.hpp -- 2-D/3-D FDM stencil generators (reference: tools/matrix_generation.hpp:47-88 builds
// the 2-D Laplacian through a std::map on the host; here the CSR arrays are written on the device, closed-form row offsets).
#ifndef VIENNACL_B200_TOOLS_MATRIX_GENERATION_HPP
#define VIENNACL_B200_TOOLS_MATRIX_GENERATION_HPP
#include "viennacl/compressed_matrix.hpp"
namespace viennacl
{
namespace tools
{
  /** @brief 5-/7-point stencil with first-order upwind convection c (c = 0: Laplacian); points_z == 1 selects 2-D */
  template<typename NumericT, unsigned int AlignmentV>
  void generate_fdm_stencil(viennacl::compressed_matrix<NumericT, AlignmentV> & A, vcl_size_t points_x, vcl_size_t points_y, vcl_size_t points_z = 1,
                            double cx_ = 0, double cy_ = 0, double cz_ = 0)
  {
    const NumericT cx = NumericT(cx_), cy = NumericT(cy_), cz = NumericT(cz_);
    ViennaCLBackend b = viennacl::backend::b200::handle();
    long long rows = 0, nnz = 0;
    viennacl::backend::b200::check(viennacl::backend::b200::abi<NumericT>::generate_stencil(b, ViennaCLInt(points_x), ViennaCLInt(points_y), ViennaCLInt(points_z), cx, cy, cz,
                                                                 NULL, NULL, NULL, &rows, &nnz));
    A = viennacl::compressed_matrix<NumericT, AlignmentV>(vcl_size_t(rows), vcl_size_t(rows), vcl_size_t(nnz));
    viennacl::backend::b200::check(viennacl::backend::b200::abi<NumericT>::generate_stencil(b, ViennaCLInt(points_x), ViennaCLInt(points_y), ViennaCLInt(points_z), cx, cy, cz,
                                                                 A.handle1().template ptr<unsigned int>(), A.handle2().template ptr<unsigned int>(),
                                                                 A.handle().template ptr<NumericT>(), &rows, &nnz));
    A.generate_row_block_information();
  }

  /** @brief The reference's generator name and meaning: 2-D Laplace, diagonal 4, neighbours -1 */
  template<typename NumericT, unsigned int AlignmentV>
  void generate_fdm_laplace(viennacl::compressed_matrix<NumericT, AlignmentV> & A, vcl_size_t points_x, vcl_size_t points_y)
  { generate_fdm_stencil(A, points_x, points_y, 1); }
}
}
#endif
