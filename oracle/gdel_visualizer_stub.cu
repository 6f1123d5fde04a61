// oracle/gdel_visualizer_stub.cu — no-op replacement of the reference's OpenGL viewer (source/gDel2D/Visualizer.cpp,
// 1138 lines of GLUT code), so that the UNMODIFIED gDel2D sources link without OpenGL.  TEST INFRASTRUCTURE ONLY.
// GpuDelaunay.cu calls Visualizer::instance()->isEnable() / addFrame(...) at :256, :666-706, :975, :1224; every one of
// them does nothing here.  The class declaration is the reference's own header.
#include "Visualizer.h"

Visualizer *Visualizer::_singleton = NULL;

Visualizer *Visualizer::instance() {
    // The reference's constructor is private and sets up GL state; the stub never dereferences `this`.
    if (_singleton == NULL) _singleton = (Visualizer *)::operator new(sizeof(Visualizer));
    return _singleton;
}
bool Visualizer::isEnable() { return false; }
void Visualizer::addFrame(const Point2HVec &, const SegmentHVec &, const TriHVec &, const IntHVec &, int) {}
void Visualizer::addFrame(const Point2HVec &, const SegmentHVec &, const TriHVec &, int) {}
void Visualizer::addFrame(const Point2DVec &, const TriDVec &, int) {}
void Visualizer::addFrame(const Point2DVec &, const SegmentDVec &, const TriDVec &, int) {}
void Visualizer::addFrame(const Point2DVec &, const SegmentDVec &, const TriDVec &, const IntHVec &, int) {}
