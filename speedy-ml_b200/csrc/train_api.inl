// train_api.inl -- C entry points of the training path (included at the end of engine.cu).
extern "C" {

int sml_train_begin(sml_engine *h, int, const int32_t *, int, int) { FAIL(h, "training path not built yet"); }
int sml_train_feed(sml_engine *h, const double *, const int64_t *, const double *, const int64_t *, int, int) { FAIL(h, "training path not built yet"); }
int sml_train_solve(sml_engine *h, double, double, int, double, int32_t *) { FAIL(h, "training path not built yet"); }
int sml_train_gram_get(sml_engine *h, int, double *, double *) { FAIL(h, "training path not built yet"); }
int sml_train_end(sml_engine *h) { FAIL(h, "training path not built yet"); }
int sml_mldivide(sml_engine *h, double *, int, double *, int, int, int) { FAIL(h, "training path not built yet"); }

}  // extern "C"
