// dic_types.hpp -- the vocabulary of the reference's engine API (interface only):
// enums.hpp:10-109 and domains.hpp:10-118 of namascar/correlation, with CorrelationResult widened
// to 12 parameters (the quadratic extension) and per-level accounting. A translation unit that
// already includes the reference's own enums.hpp / domains.hpp defines DIC_USE_REFERENCE_TYPES and
// gets none of this.
#pragma once
#include <utility>
#include <vector>

#include "../../include/dic_b200.h"

#ifndef DIC_USE_REFERENCE_TYPES
enum interpolationModelEnum { im_nearest, im_bilinear, im_bicubic, im_NUMBER_OF_ITEMS };
enum fittingModelEnum { fm_U, fm_UV, fm_UVQ, fm_UVUxUyVxVy, fm_UVUxUyVxVyQuad /* extension */, fm_NUMBER_OF_ITEMS };
enum errorEnum {
  error_none, error_model_out_of_image, error_interpolation_out_of_image,
  error_correlation_max_iters_reached, error_bad_domain, error_cuSolver, error_cuda,
  error_multiThread, error_NUMBER_OF_ITEMS
};
enum colorEnum { color_monochrome, color_color, color_NUMBER_OF_ITEMS };
enum deformationDescriptionEnum { def_strict_Lagrangian, def_Lagrangian, def_Eulerian, def_NUMBER_OF_ITEMS };
enum errorHandlingModeEnum { errorMode_stopAll, errorMode_stopFrame, errorMode_continue, errorMode_NUMBER_OF_ITEMS };
enum referenceImageEnum { refImage_First, refImage_Previous, refImage_NUMBER_OF_ITEMS };
typedef std::vector<std::pair<float, float>> v_points;
struct frame_results; // the manager's bookkeeping record; correlate() never reads it (cuda_class.cu:104-293)
#endif

// domains.hpp:110-118, widened
struct CorrelationResult {
  float resultingParameters[DIC_MAX_PARAMS];
  float chi;
  int numberOfPoints;
  int iterations;
  errorEnum errorCode{error_none};
  float undCenterX;
  float undCenterY;
  int iterationsPerLevel[DIC_MAX_LEVELS];
  int evaluationsPerLevel[DIC_MAX_LEVELS];
  int pointsPerLevel[DIC_MAX_LEVELS];
};
static_assert(sizeof(CorrelationResult) == sizeof(dic_result), "CorrelationResult must mirror dic_result");
