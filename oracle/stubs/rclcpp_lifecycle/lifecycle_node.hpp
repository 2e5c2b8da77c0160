// Oracle build shim: only the SharedPtr typedef is referenced (Mapper.h:982).
#pragma once
#include <memory>
namespace rclcpp_lifecycle {
class LifecycleNode { public: typedef std::shared_ptr<LifecycleNode> SharedPtr; };
}  // namespace rclcpp_lifecycle
