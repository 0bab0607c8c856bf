// rttechnique.h -- the drop-in boundary type (reflectcuts/realtimetechniques/rttechnique.h:6-10).
#pragma once
#include "rtcommon.h"

namespace evplp_host {

class RtTechnique {
public:
    virtual ~RtTechnique() {}
    virtual void render(shared_ptr<RtScene>& scene, const Vec2& resolution, const Json& json) = 0;
};

}  // namespace evplp_host
