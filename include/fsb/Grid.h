// GridInterface and BBox of the reference (include/Grid.h:11-63,186-190).  The storage class
// Grid<T> itself does not exist on this side of the boundary: every grid lives in HBM inside the
// fsb context and is reached through MacGrid's accessors.
#ifndef FSB_GRID_H
#define FSB_GRID_H

#include <cassert>

#include "MathDefinitions.h"

class GridInterface
{
public:
  GridInterface(int size_x, int size_y, MyFloat delta_x = 1, MyFloat delta_y = 1)
      : _SIZE_X(size_x), _SIZE_Y(size_y), _DELTA_X(delta_x), _DELTA_Y(delta_y) {}

  void linearTo2D(int idx, int* i, int* j) const { *i = idx % _SIZE_X; *j = idx / _SIZE_X; }
  int twoDToLinear(int i, int j) const { assert(indexIsValid(i, j)); return i + j * _SIZE_X; }
  void worldToCell(MyFloat x, MyFloat y, int* i, int* j) const
  {
    *i = (int)(x / _DELTA_X); // truncation toward zero, as upstream
    *j = (int)(y / _DELTA_Y);
  }
  void cellToWorld(int i, int j, MyFloat* x, MyFloat* y) const
  {
    assert(indexIsValid(i, j));
    *x = i * _DELTA_X;
    *y = j * _DELTA_Y;
  }
  bool indexIsValid(int i, int j) const { return i >= 0 && i < _SIZE_X && j >= 0 && j < _SIZE_Y; }

  int sizeX() const { return _SIZE_X; }
  int sizeY() const { return _SIZE_Y; }
  MyFloat deltaX() const { return _DELTA_X; }
  MyFloat deltaY() const { return _DELTA_Y; }
  MyFloat lengthX() const { return _SIZE_X * _DELTA_X; }
  MyFloat lengthY() const { return _SIZE_Y * _DELTA_Y; }

protected:
  int _SIZE_X, _SIZE_Y;
  MyFloat _DELTA_X, _DELTA_Y;
};

template <class T>
struct BBox
{
  T x_min, x_max, y_min, y_max;
};

#endif
